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

#ifdef __cplusplus
}
#endif
#endif /* SGS_H */
