/* libsgs - B200 (sm_100a) kernels for the sEEG -> audio decoding path of
 * cognitive-systems-lab/closed-loop-seeg-speech-synthesis, behind a plain C ABI.
 *
 * The reference is pure Python; its "FFI" for this path is the set of numpy/scipy/sklearn calls its
 * nodes make.  Each entry point below names the reference call sites it replaces (paths relative to
 * the reference tree).  The Python host code that mirrors the reference's Node / function API
 * (closed-loop-seeg-speech-synthesis_b200/{livenodes,local,train.py,decode.py}) binds these symbols with
 * ctypes (sgs/_lib.py); INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; sgs_last_error() describes the failure.
 *   - data pointers may be HOST or DEVICE memory (decided per pointer with cudaPointerGetAttributes);
 *     host buffers are copied on `stream` inside the call.  Small tables (coefficients, window starts,
 *     weights) are always host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous
 *     with respect to device buffers and synchronous (stream-synchronised before return) when any
 *     output buffer is host memory.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SGS_H
#define SGS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SGS_ABI_VERSION 1

int sgs_abi_version(void);
const char* sgs_last_error(void);
/* Selects the device for the calling thread and creates the context (lazily, so it is safe to load the
 * library before fork() and call this in the child: livenodes/Sender.py:57-64 forks the whole graph). */
int sgs_init(int device);
int sgs_device_count(int* count);
int sgs_synchronize(void* stream);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
unsigned long long sgs_launch_count(void);
/* Per-kernel-class device time, measured with CUDA events on the launching stream (bench.py's roofline leg).
 * Enable, run, then read: name in {iir_init, iir_state, iir_carry, iir_feat, iir_pieces_tail, iir_pieces_state, iir_pieces_feat, stack, lda,
 * lda_pack, lda_tc, gl_blocks, gl_ola, lowpass, stream, gl_batch, logmel, train, train_tc}.  iir_state / iir_feat are the
 * (stream group x time chunk) grid of the feature scan, iir_pieces_* its balanced-pieces form (large jobs; _tail = the modal sums of sgs_feat_plan_set_tail): the launch count
 * of a class tells which decomposition a call took. */
int sgs_profile_enable(int on);
int sgs_profile_read(const char* name, double* total_ms, unsigned long long* launches);

/* ---------------------------------------------------------------------------------------------------
 * High-gamma feature extraction.
 * Replaces: scipy.signal.sosfilt x3 + window energy/log  (local/offline.py:31-109,
 *           livenodes/FrameBuffer.py:139-143, livenodes/ECogFeatCalc.py:118-124).
 * ------------------------------------------------------------------------------------------------- */
typedef struct sgs_feat_plan sgs_feat_plan;

/* n_filters: 3 (50 Hz: band-pass + 2 notches) or 2 (60 Hz).  coef[n_filters*8][5] = b0 b1 b2 a1 a2 per
 * biquad (a0 = 1).  zi_unit[n_filters*8][2] = scipy.signal.sosfilt_zi of each filter.  zi_last_warm[8][2] =
 * state of the last filter after its warm-start zero fill; zero_fill_response[zero_fill] = what it emitted
 * there (online framing covers it, FrameBuffer.py:95-98). */
int sgs_feat_plan_create(sgs_feat_plan** plan, int n_filters, const double* coef, const double* zi_unit,
                         const double* zi_last_warm, const double* zero_fill_response, int zero_fill);
void sgs_feat_plan_destroy(sgs_feat_plan* plan);

/* Modal tail of the warm-up (optional; sgs/modal.py derives it from the same coefficients).  A time piece of a large job
 * starts from the state the cascade had there; without a tail that state comes from running all sections over `horizon`
 * samples from zero.  With it, only the last near_len samples go through the cascade, started from
 * state_matrix (Re c, Im c) with c[m] = sum_k lambda[m]^k x[t - 1 - k] over k < mode_len[m] - one complex geometric sum for
 * each of the n_modes (4, 8, 12 or 16; 0 removes the tail) pole pairs that outlive near_len samples.  Restates what
 * scipy.signal.sosfilt's state holds after a long run (local/offline.py:63-97) to the tolerance the tables were built for.
 * mode_len[n_modes]: multiples of 32, longest first, equal within groups of 4.  lam[n_modes][2] = lambda (re, im; a padding
 * mode is 0).  warp_blocks[16][4][2] = the blocks [lo, hi) of 32 samples (block 0 ends at t - near_len) of each mode group
 * that each of the kernel's 16 warps sums - every block of a group exactly once, a warp's blocks of one group contiguous;
 * warp_shift[16][n_modes][2] = lambda^(32 lo) of the mode's group.  state_matrix[2 n_biquads][2 n_modes], row-major.
 * kappa[n_modes][2][2 n_biquads] (re row, im row): what a state s leaves in mode m, lambda^n (kappa[m] . s) after n samples -
 * a piece that starts nearer to the beginning of the recording than the longest mode_len sums what samples there are and
 * takes the rest from the reference's initial state (local/offline.py:39-62). */
int sgs_feat_plan_set_tail(sgs_feat_plan* plan, int near_len, int n_modes, const int32_t* mode_len, const double* lam,
                           const int32_t* warp_blocks, const double* warp_shift, const double* state_matrix,
                           const double* kappa);

/* x: [n_sessions][n_samples][n_channels] (fp32 if x_is_f64 == 0, else fp64), sessions `session_stride`
 *    elements apart (0 = dense).
 * win_starts[n_windows]: first sample of each analysis window (strictly increasing; may be negative down to
 *    -zero_fill for the online framing), every window is `window_len` samples long and must end <= n_samples.
 * feat: [n_sessions][n_windows][n_channels] fp64 = log(sum(y^2) + 0.01).
 * Time is cut into n_chunks chunks of chunk_len samples (the last takes the remainder); `horizon` is the
 * zero-state pass length; phi = A^chunk_len ((2*nb) x (2*nb), row-major) must be given when
 * horizon >= chunk_len and n_chunks > 2, else NULL. */
int sgs_feat_extract(sgs_feat_plan* plan, const void* x, int x_is_f64, int64_t n_samples, int n_channels,
                     int n_sessions, int64_t session_stride, const int32_t* win_starts, int n_windows,
                     int window_len, int n_chunks, int64_t chunk_len, int horizon, const double* phi,
                     double* feat, void* stream);

/* Temporal context stacking (local/offline.py:111-116, ECogFeatCalc.py:137-144):
 * out[s][r][c*(order+1)+tap] = feat[s][r + first_row - (order-tap)*step][c], zero where the index is < 0.
 * offline: first_row = order*step, n_rows = n_windows - order*step; online: first_row = 0, n_rows = n_windows. */
int sgs_feat_stack(const double* feat, int n_sessions, int n_windows, int n_channels, int n_rows, int first_row,
                   int order, int step, double* out, void* stream);

/* Streaming form (the ECogFeatCalc node: livenodes/ECogFeatCalc.py:67-104 over livenodes/FrameBuffer.py:60-177).
 * All filter state, the y^2 history and the 21-row stack buffer stay on the device between pushes.
 * x: n x n_channels new samples (host, n <= 128); frame_ends[n_frames]: exclusive end (in real-sample
 * coordinates; the warm-start zero fill is negative) of every frame this push completes and frame_index[] their
 * running numbers - the host keeps the reference's fractional frame schedule (FrameBuffer.py:177);
 * out: n_frames x (n_channels*(order+1)) stacked rows (host).  Synchronous when n_frames > 0. */
typedef struct sgs_feat_stream sgs_feat_stream;
int sgs_feat_stream_create(sgs_feat_stream** stream_out, sgs_feat_plan* plan, int n_channels, int frame_size, int order,
                           int step);
void sgs_feat_stream_destroy(sgs_feat_stream* s);
int sgs_feat_stream_push(sgs_feat_stream* s, const void* x, int x_is_f64, int n, const int64_t* frame_ends,
                         const int64_t* frame_index, int n_frames, double* out, void* stream);

/* ECogFeatCalc(warm_start=False), FrameBuffer.py:91-98: the last filter starts from zi * (its first input) like the others
 * instead of its warm state, and there is no zero fill in front of the stream (frame_ends are then plain sample counts; the
 * caller also withholds the first order * step stacked rows, as the empty stack FrameBuffer does).  Before the first push. */
int sgs_feat_stream_set_cold_start(sgs_feat_stream* stream, int cold);

/* Filtering FrameBuffer node (livenodes/FrameBuffer.py:86-143): one scipy.signal.sosfilt cascade (<= 8 sections) over all
 * channels with the state resident between chunks.  sos[n_sections][6], zi[n_sections][2] = sosfilt_zi(sos); the first push
 * starts from zi (warm_start, FrameBuffer.py:95-98 then pushes its zero fill through the same stream) or from zi * x[0]
 * (cold start, FrameBuffer.py:90-92).  y[n][n_channels] float64 = the filtered chunk.  Synchronous. */
typedef struct sgs_sos_stream sgs_sos_stream;
int sgs_sos_stream_create(sgs_sos_stream** stream_out, const double* sos, const double* zi, int n_sections, int n_channels,
                          int warm_start);
void sgs_sos_stream_destroy(sgs_sos_stream* s);
int sgs_sos_stream_push(sgs_sos_stream* s, const void* x, int x_is_f64, int n, double* y, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Per-mel-bin LDA decoding + dequantisation (+ optional smoothing across bins).
 * Replaces: 40 x sklearn LinearDiscriminantAnalysis.predict per frame (livenodes/LDASynthesis.py:25-26),
 *           medians lookup + scipy.ndimage.gaussian_filter (livenodes/Dequantization.py:16-17),
 *           local/quantization.py:125-135 (dequantize_spectrogram, no smoothing).
 * ------------------------------------------------------------------------------------------------- */
typedef struct sgs_lda_model sgs_lda_model;

/* W[n_bins][n_classes][n_features], bias[n_bins][n_classes] (-inf for classes a bin never saw),
 * class_labels[n_bins][n_classes] (the label value each score row stands for), select[n_features] = column of the
 * stacked feature vector (c*(order+1)+tap) each model feature reads, medians[n_bins][n_levels],
 * smooth_taps[2*smooth_radius+1] (NULL / 0 = model cannot smooth).  Binary estimators are packed by the caller
 * as the two rows {0, coef_} / {0, intercept_}. */
int sgs_lda_model_create(sgs_lda_model** model, int n_bins, int n_classes, int n_features, const double* W,
                         const double* bias, const double* class_labels, const int32_t* select, const double* medians,
                         int n_levels, const double* smooth_taps, int smooth_radius);
void sgs_lda_model_destroy(sgs_lda_model* model);

/* feat: un-stacked log-power [n_sessions][n_windows][n_channels] (output of sgs_feat_extract); row r of the
 * stacked view is assembled on the fly exactly as sgs_feat_stack would (first_row/order/step as there; pass
 * order = 0, first_row = 0 to decode already-stacked rows of width n_channels).
 * labels, spec: [n_sessions][n_rows][n_bins] fp64, either may be NULL.  smooth != 0 applies the taps across bins
 * with scipy's 'reflect' boundary.
 * A model handle is SINGLE-STREAM: the call keeps per-call scratch (centring vector, re-score list, its counter) in the
 * handle, so two decodes of one model must not be in flight on two streams at once - create one model per stream. */
int sgs_lda_decode(const sgs_lda_model* model, const double* feat, int n_sessions, int n_windows, int n_channels,
                   int n_rows, int first_row, int order, int step, double* labels, double* spec, int smooth,
                   void* stream);
/* For batches of >= 4096 frames sgs_lda_decode scores on the tensor cores (split-TF32 tcgen05 GEMM) and re-scores in
 * fp64 every (frame, bin) whose top-two gap is below the tensor-core error bound; this returns how many (frame, bin)
 * pairs - of frames x n_bins - the last such call re-scored.  SGS_LDA_TC=0 in the environment forces the pure fp64 kernel. */
int sgs_lda_last_rescored(const sgs_lda_model* model, int* n_pairs);

/* ---------------------------------------------------------------------------------------------------
 * Griffin-Lim, streaming-node semantics (livenodes/GriffinLim.py:13-174), batched over frames and sessions.
 * Replaces: numpy.fft.rfft/irfft + np.angle/np.exp per frame (GriffinLim.py:64-96), the ring-buffer overlap-add
 *           (GriffinLim.py:145-166), scipy.signal.lfilter + int16 conversion (GriffinLim.py:169-174).
 * ------------------------------------------------------------------------------------------------- */
typedef struct sgs_gl_node sgs_gl_node;

/* window = blackman(fft_size); ola_window = blackman(block_len*hop); inv_idx/inv_w[bins][2] = the (at most two)
 * non-zero entries of MelFilterBank.melInvMatrix per spectral bin; lp_b/lp_a[lp_order+1] = the output low-pass;
 * lp_phi[lp_order^2] = (zero-input DF2T state transition)^lp_chunk and lp_phi_sub = the same ^64 for the two-level
 * chunked scan (lp_chunk must be 2048); norm_div = normFactor*1.01.
 * Only the configuration decode.py uses is built: fft 256, hop 160, block_len 3, context_width 1. */
int sgs_gl_node_create(sgs_gl_node** node, int fft_size, int hop, int block_len, int context_width, int n_mels,
                       const double* window, const double* ola_window, const int32_t* inv_idx, const double* inv_w,
                       const double* lp_b, const double* lp_a, int lp_order, const double* lp_phi, int lp_chunk,
                       const double* lp_phi_sub, double norm_div, int iterations);
void sgs_gl_node_destroy(sgs_gl_node* node);

/* logmel[n_sessions][n_frames][n_mels]; positions[n_frames] = write-head position after each frame (host table:
 * int((ms/1000)*sampleRate) with ms accumulated as the node does); noise[n_sessions][n_frames][480] = the
 * np.random.rand(480) draw of each frame (row 0 unused) or NULL to draw from the counter-based generator with
 * `seed`.  lp_state[n_sessions][lp_order] (host, in/out, NULL = start from rest).  Outputs per session:
 * n_out = positions[n_frames-1] - positions[0] samples: pcm int16, optionally the un-quantised low-passed signal
 * (`filtered`) and the raw 480-sample blocks (`blocks_out`, [n_sessions][n_frames][480]). */
int sgs_gl_node_synthesize(sgs_gl_node* node, const double* logmel, int n_sessions, int n_frames,
                           const int32_t* positions, const double* noise, uint64_t seed, double* lp_state,
                           int16_t* pcm, double* filtered, double* blocks_out, void* stream);

/* Streaming form (GriffinLimSynthesis.add_data): feed n (1..16) new spectral frames, receive the audio the node
 * emits for them.  pos[n] = write head after each new frame, pos_before = write head before the first of them;
 * noise[n][480] or NULL (counter-based generator with `seed`).  The previous spectral frame, the last 32 blocks and
 * the low-pass state stay on the device.  pcm must hold sum(pos[i]-pos[i-1]) samples; *n_pcm receives the count
 * (0 for the very first frame, GriffinLim.py:131-132).  Synchronous. */
int sgs_gl_node_push(sgs_gl_node* node, const double* logmel, int n, const int32_t* pos, int32_t pos_before,
                     const double* noise, uint64_t seed, int16_t* pcm, int* n_pcm, void* stream);

/* Write-head positions are absolute sample counts in 32 bits (37 h of audio at 16 kHz); only their differences matter.
 * Subtracts `delta` from every position the node remembers, so that a long-running caller can keep passing small numbers
 * (the reference keeps its positions modulo the ring length, GriffinLim.py:115-166, and so runs indefinitely): after the
 * call, pos / pos_before of sgs_gl_node_push and sgs_chain_push are expected in the shifted coordinates. */
int sgs_gl_node_rebase(sgs_gl_node* node, int32_t delta);

/* GriffinLimSynthesis(useLogMels=...), GriffinLim.py:84-87: 1 (default) takes log-mel frames through fromLogMels (exp, then
 * non-finite magnitudes -> 0), 0 takes linear mel frames through fromMels.  Applies to the calls that follow. */
int sgs_gl_node_set_log_mels(sgs_gl_node* node, int log_mels);

/* ---------------------------------------------------------------------------------------------------
 * Fused streaming chain: the four nodes decode.py:152-183 wires (ECogFeatCalc -> LDASynthesis -> Dequantization ->
 * GriffinLimSynthesis), driven back to back on one CUDA stream with ONE host<->device round trip per packet of
 * samples (the 10 ms-frame latency path).  The chain borrows the three streaming handles: their state is the same
 * state the single-node entry points above advance, so a chain push equals sgs_feat_stream_push + per frame
 * sgs_lda_decode(smooth) + sgs_gl_node_push, bit for bit.
 *   x[n][n_channels] host samples (n <= 128); frame_ends / frame_index[n_frames] as for sgs_feat_stream_push;
 *   gl_pos[n_frames] / gl_pos_before / noise[n_frames][480] / seed as for sgs_gl_node_push.
 *   rows[n_frames][n_channels*(order+1)], labels[n_frames][n_bins], spec[n_frames][n_bins] (smoothed),
 *   pcm[sum of hops] and *n_pcm receive what the four nodes emit.  Synchronous when n_frames > 0.
 * ------------------------------------------------------------------------------------------------- */
typedef struct sgs_chain sgs_chain;
int sgs_chain_create(sgs_chain** chain, sgs_feat_stream* feat, int n_channels, const sgs_lda_model* lda, sgs_gl_node* gl);
void sgs_chain_destroy(sgs_chain* chain);
int sgs_chain_push(sgs_chain* chain, const void* x, int x_is_f64, int n, const int64_t* frame_ends, const int64_t* frame_index,
                   int n_frames, const int32_t* gl_pos, int32_t gl_pos_before, const double* noise, uint64_t seed,
                   double* rows, double* labels, double* spec, int16_t* pcm, int* n_pcm, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Batch Griffin-Lim (local/offline.py:131-192): 50 ms periodic-Hann windows, 10 ms hop, complex phase projection.
 * window[800]; inv_idx/inv_w[401][2] = 2-tap form of MelFilterBank(401, n_mels, 16000).melInvMatrix.
 * logmel[n_utt][n_frames][n_mels]; noise[n_utt] rows `noise_stride` doubles apart, each >= 160*(n_frames-1)+800
 * samples = the head of the reference's np.random.rand(2*T*401) start.  pcm[n_utt][160*n_frames] =
 * int16(x / max|x| * 32767); waveform (optional) = the un-scaled float result.
 * ------------------------------------------------------------------------------------------------- */
typedef struct sgs_gl_batch sgs_gl_batch;
int sgs_gl_batch_create(sgs_gl_batch** plan, int win_len, int hop, int n_mels, const double* window,
                        const int32_t* inv_idx, const double* inv_w);
void sgs_gl_batch_destroy(sgs_gl_batch* plan);
int sgs_gl_batch_synthesize(sgs_gl_batch* plan, const double* logmel, int n_utt, int n_frames, const double* noise,
                            int64_t noise_stride, int iterations, int16_t* pcm, double* waveform, void* stream);

/* Audio -> log-mel target (local/offline.py:219-241 as train.py:128 calls it: 16 ms symmetric-Hann windows, 10 ms
 * shift, win_len - shift zeros in front): out[n_frames][n_mels] = log(|rfft(window * frame)| . mel + 1e-7).
 * mel[n_bins][n_mels] = MelFilterBank.melMatrix. */
int sgs_logmel(const double* audio, int64_t n_audio, const double* window, int win_len, int shift, const double* mel,
               int n_bins, int n_mels, int64_t n_frames, double* out, void* stream);

/* Test hook: out[i] = exp(np.angle(re[i] + 1j*im[i])), the per-bin operation of the node's phase step (GriffinLim.py:93)
 * exactly as k_gl_blocks evaluates it (csrc/exp_angle.cuh). */
int sgs_exp_angle(const double* im, const double* re, int64_t n, double* out, void* stream);

/* scipy.signal.decimate(audio, q) as train.py:125 uses it (ftype='iir', zero_phase=True): out[j] = sosfiltfilt(sos, audio)[q*j],
 * j < ceil(n/q).  sos[n_sections][6] = cheby1(8, 0.05, 0.8/q, 'sos'), zi[n_sections][2] = sosfilt_zi(sos),
 * edge = 3*(2*n_sections+1) (sosfiltfilt's odd-extension length), warm = samples after which the filter state has
 * forgotten its start to 2^-70 (host: from the largest pole radius).  Both passes run as chunked scans over time. */
int sgs_decimate(const double* audio, int64_t n, int q, const double* sos, const double* zi, int n_sections, int edge, int warm,
                 double* out, void* stream);

/* Dequantization node alone (livenodes/Dequantization.py:15-18; local/quantization.py:125-135 with smooth = 0):
 * out[r][b] = medians[b][labels[r][b]], optionally smoothed across bins with taps[2*radius+1] ('reflect'). */
int sgs_dequantize(const double* medians, int n_bins, int n_levels, const double* taps, int radius,
                   const double* labels, int64_t n_rows, int smooth, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Training side (train.py:78-168).
 * ------------------------------------------------------------------------------------------------- */
/* Per-column min / max of y[n][ncol] (the logistic borders of local/quantization.py:83-109 are derived from them). */
int sgs_col_minmax(const double* y, int64_t n, int ncol, double* mn, double* mx, void* stream);
/* quantize_spectrogram (local/quantization.py:112-122): labels[r][c] = smallest k with y[r][c] <= borders[c][k], else 0. */
int sgs_quantize(const double* y, int64_t n, int ncol, const double* borders, int n_intervals, double* labels, void* stream);
/* feature_selection's statistic (train.py:96-109): rho[c] = scipy.stats.spearmanr(x[:, c], mean(y, axis=1)) with
 * average ranks for ties; colsum[c] = sum(x[:, c]) (the caller zeroes rho where that is ~0).  x rows are row_stride
 * doubles apart (0 = ncol). */
int sgs_spearman(const double* x, int64_t n, int ncol, int64_t row_stride, const double* y, int ny, double* rho,
                 double* colsum, void* stream);
/* Mean of the selected columns (global centring of the LDA statistics). */
int sgs_col_means(const double* x, int64_t n, int64_t row_stride, const int32_t* select, int n_features, double* xbar,
                  void* stream);
/* Sufficient statistics of LinearDiscriminantAnalysis(solver='svd').fit for all bins at once (train.py:112-118):
 * with Xc = x[:, select] - xbar: G = Xc^T Xc [F][F], class_sums[bin][class][F] = sum of Xc rows per label,
 * counts[bin][class].  xbar_in: centre on this vector (multi-GPU: the all-reduced global mean) or NULL to use the
 * mean of the given rows; xbar receives the vector used.  All four outputs are additive across row shards that
 * share xbar_in, which is what the NCCL all-reduce of train.py's multi-GPU path sums. */
int sgs_lda_stats(const double* x, int64_t n, int64_t row_stride, const int32_t* select, int n_features,
                  const double* labels, int n_bins, int n_classes, const double* xbar_in, double* xbar, double* G,
                  double* class_sums, double* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGS_H */
