// Internal interface between the kernel files and the C-ABI files of libsgs: every structure that crosses a translation
// unit (kernel parameter blocks, device-table bundles) and every host-side launcher is declared HERE and only here.
#pragma once
#include "common.cuh"
#include "feat.cuh"
#include "fft.cuh"

struct sgs_feat_stream;
struct sgs_lda_model;
struct sgs_gl_node;

namespace sgs {

// ---- feature extraction (feat.cu, stream.cu) --------------------------------------------------------------------------------
int feat_run(int n_biquads, bool monic, const void* x, bool x_is_f64, double* feat, double* slots, const double* phi,
             bool apply_phi, const long long* bounds, const int* kfirst, const int* starts, const double* zf,
             const double* coef, const FeatCoefs& cf, const FeatGeom& g, cudaStream_t st);
int feat_run_pieces(int n_biquads, bool monic, const void* x, bool x_is_f64, double* feat, double* init_state, double* seg_state,
                    const FeatSeg* segs, const int* piece_first, int n_pieces, int n_segs, const TailTab* tail /* or nullptr */,
                    const double* tail_matrix /* device, [2 nb][2 n_modes] */, const int* starts, const double* zf,
                    const FeatCoefs& cf, const FeatGeom& g, cudaStream_t st);
int stack_run(const double* feat, double* out, int n_sessions, int n_windows, int n_channels, int n_rows, int first_row,
              int order, int step, cudaStream_t st);

constexpr int kSqRing = 256;            // per-channel history of y^2 (>= frame_size + largest sub-chunk)
constexpr int kFeatRing = 32;           // per-channel history of log-power rows (>= order*step + 1)
constexpr int kMaxFramesPerPush = 16;
struct StreamFrames {                   // frames completed by one streaming push (the host computes the schedule)
    int n;
    long long end[kMaxFramesPerPush];   // exclusive end, in real-sample coordinates (zero fill = negative)
    long long index[kMaxFramesPerPush]; // running frame number k
};
int feat_stream_run(int n_biquads, const void* x, bool x_is_f64, int n, int n_channels, long long t0, double* z,
                    double* sq_ring, double* feat_ring, const double* zf, int zero_fill, int cold_last, int frame_size, int order,
                    int step, double* out, const FeatCoefs& cf, const StreamFrames& fr, cudaStream_t st);

constexpr int kSosMaxSections = 8;
struct SosCoefs { double c[kSosMaxSections][5]; double zi[kSosMaxSections][2]; int n_sections; };   // b0 b1 b2 a1 a2; sosfilt_zi
int sos_stream_run(const void* x, bool x_is_f64, int n, int n_channels, double* z, double* y, int first, int warm_start,
                   const SosCoefs& cf, cudaStream_t st);

// ---- LDA decode (lda.cu, lda_tc.cu) and dequantisation (stream.cu) ---------------------------------------------------------------
struct LdaGeom {
    int n_bins, n_classes, n_features, n_levels;
    int n_windows, n_channels, n_rows, first_row, order, step;
    int smooth_radius;
};
int lda_run(const double* feat, const double* Wt, const double* bias, const double* cls, const int* select,
            const double* medians, const double* taps, double* labels, double* spec, int smooth, int n_sessions,
            const LdaGeom& g, cudaStream_t st, const int* list, const int* list_count, long long list_cap);
int lda_pairs_run(const double* feat, const double* Wt, const double* bias, const double* cls, const int* select, double* labels,
                  const LdaGeom& g, cudaStream_t st, const int* flags, const int* list, const int* list_count, const int* slice_bins,
                  long long list_cap);
struct LdaTcGeom {
    int n_windows, n_channels, n_rows, first_row, order, step, n_bins, n_features;
    int tiles_per_session, n_tiles;
    double eps;                         // relative error bound of the tensor-core score
};
int lda_tc_run(const double* feat, const float* Bmat, const double* Wt, const double* bias0, const double* chan_mean, double* bias,
               const double* cls, const int* feat_chan, const int* feat_back, double* centre, const int* slice_bins, const double* wnorm,
               double* labels, int* flags, int* list, int* count, long long n_frames_total, const LdaTcGeom& g, cudaStream_t st);
int dequantize_run(const double* labels, const double* medians, const double* taps, int radius, int smooth, int n_bins,
                   int n_levels, long long n_rows, double* out, cudaStream_t st);

// ---- Griffin-Lim, node semantics (gl_node.cu, gl_blocks8.cuh, stream.cu) ----------------------------------------------------------
constexpr int kFft = 256, kHalf = 128, kHop = 160, kBlk = 480, kBins = 129;
constexpr int kLpMaxOrd = 8;
constexpr int kBlockRing = 32;          // >= kMaxFramesPerPush + 4: a push writes all its blocks before the first hop is emitted
struct GlNodeTables {                   // device pointers, built once per node configuration
    const double* window;               // blackman(256)
    const cplx* tw_full;                // exp(-2 pi i k / 256), k <= 128
    const cplx* tw_t;                   // [2][16][9] W128^((l + 48 f) k1) (register-FFT kernel, gl_blocks8.cuh)
    const int* inv_idx;                 // [129][2] mel index of each inverse-mel tap
    const double* inv_w;                // [129][2] weight (0 where unused)
    int log_mels;                       // 1: input frames are log-mels (fromLogMels), 0: linear mels (fromMels, GriffinLim.py:84-87)
};
struct LpCoefs { double b[kLpMaxOrd + 1], a[kLpMaxOrd + 1]; int ord; };
struct EmitFrames {                     // streaming emission schedule of one push
    int n;
    long long index[kMaxFramesPerPush]; // frame number k
    int pos[kMaxFramesPerPush];         // write head after frame k
    int prev[kMaxFramesPerPush];        // write head before frame k
    int ring_pos[kBlockRing];           // write head of the block stored in each ring slot (after this push)
    long long ring_index[kBlockRing];   // its frame number (-1 = empty)
};
int gl_blocks_run(const double* logmel, const double* noise, unsigned long long seed, double* blocks, const GlNodeTables& tab,
                  int n_sessions, int n_frames, int n_mels, int first_frame, int iters, long long ring_base, int ring_len,
                  cudaStream_t st);
int gl_emit_run(const double* blocks, const int* pos, const double* ola_window, double* v, double* states, double* zi,
                double* zi_out, int carry_depth, const double* phi, const LpCoefs& c, double norm_div, short* pcm, double* filtered,
                int n_sessions, int n_frames, int first_frame, long long n_out, int chunk, int n_chunks, cudaStream_t st);
int gl_emit_stream_run(const double* block_ring, const double* ola_window, double* lp_state, short* pcm, const LpCoefs& c,
                       double norm_div, int first_frame, const EmitFrames& fr, cudaStream_t st);
int exp_angle_run(const double* im, const double* re, long long n, double* out, cudaStream_t st);

// ---- batch Griffin-Lim and log-mel (gl_batch.cu) ------------------------------------------------------------------------------------
struct GlBatchTables {
    const double* window;               // [800] periodic Hann
    const cplx* tw_half;                // exp(-2 pi i t / 400)
    const cplx* tw_full;                // exp(-2 pi i k / 800), k <= 400
    const int* inv_idx;                 // [401][2]
    const double* inv_w;                // [401][2]
};
int gl_batch_run(const double* logmel, double* x, const GlBatchTables& tab, int n_utt, int T, int n_mels, int iters,
                 long long x_len, double* mx, short* pcm, cudaStream_t st);
int logmel_run(const double* audio, long long n_audio, const double* window, const cplx* tw_half, const cplx* tw_full,
               const double* mel, int n_mels, long long n_frames, int shift, int pad, double* out, cudaStream_t st);

// ---- training side (train.cu, train_tc.cu) ----------------------------------------------------------------------------------------------
int quantize_run(const double* y, long long n, int ncol, const double* borders, int nint, double* q, cudaStream_t st);
int colminmax_run(const double* y, long long n, int ncol, double* mn, double* mx, cudaStream_t st);
int spearman_run(const double* x, long long n, int ncol, long long row_stride, const double* y, int ny, double* rho,
                 double* colsum, cudaStream_t st);
int col_means_run(const double* x, long long n, long long row_stride, const int* select, int nf, double* xbar, cudaStream_t st);
int lda_stats_run(const double* x, long long n, long long row_stride, const int* select, int nf, const double* labels, int n_bins,
                  int n_classes, double* xbar, double* G, double* sums, double* counts, const double* xbar_in, cudaStream_t st);
bool lda_stats_tc_supported(int nf, int n_bins, int n_classes);
int lda_stats_tc_run(const double* x, long long n, long long row_stride, const int* select, int nf, const double* labels, int n_bins,
                     int n_classes, const double* xbar, double* G, double* sums, double* counts, cudaStream_t st);

// ---- streaming handles driven back to back by the fused chain (api_feat.cu, api_lda.cu, api_gl.cu -> api_chain.cu) -----------------
int feat_stream_row_width(const sgs_feat_stream* s);
// small host-to-device copy of the streaming path.  From page-locked memory (src_pinned) a copy kernel reads the mapped
// host buffer over PCIe instead of a cudaMemcpyAsync, which keeps the packet on the compute queue (no copy-engine hand-over):
// median packet latency 0.237 -> 0.226 ms (64-sample packets), 0.190 -> 0.185 ms (32).
int copy_in_small(void* dst, const void* src, size_t bytes, bool src_pinned, cudaStream_t st);
int feat_stream_enqueue(sgs_feat_stream* s, const void* x, int x_is_f64, int n, const int64_t* frame_ends,
                        const int64_t* frame_index, int n_frames, double* d_rows, cudaStream_t st, bool src_pinned = false);
int lda_model_bins(const sgs_lda_model* m);
int lda_rows_enqueue(const sgs_lda_model* m, const double* d_rows, int n_rows, int row_width, double* d_labels, double* d_spec,
                     int smooth, cudaStream_t st);
int gl_node_enqueue(sgs_gl_node* s, const double* logmel, int n, const int32_t* pos, int32_t pos_before, const double* noise,
                    uint64_t seed, short* d_pcm, int* n_pcm, cudaStream_t st, bool src_pinned = false);

}  // namespace sgs
