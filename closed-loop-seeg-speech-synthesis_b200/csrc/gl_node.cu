// Streaming-semantics Griffin-Lim (the GriffinLimSynthesis node), batched over frames and sessions.
//
// Reference semantics restated (livenodes/GriffinLim.py, closed form R3 of SURVEY.md 8a'):
//   * per frame k >= 1 an independent 480-sample block is synthesised from the log-mels of frames k-1, k and a
//     fresh np.random.rand(480) start (GriffinLim.py:76-96): `iters` times { X_o = rfft(blackman256 * x[o:o+256])
//     for o in {0,160} (range(0, 480-256, 160), quirk Q3); Z_o = S_o * exp(angle(X_o))  -- REAL, the reference has
//     no 1j in the exponent (quirk Q1); x = sum_o irfft(Z_o) * blackman256 placed at o; x[416:480] = 0 }.
//   * S = fromLogMels: exp(logmel) through the 2-tap inverse mel matrix (local/MelFilterBank.py:82-83).
//   * ring-buffer overlap-add of the un-windowed blocks normalised by the sum of blackman(480) segments, with the
//     reference's write-head positions int((ms/1000)*16000) (GriffinLim.py:115-166; hops of 159/161 samples, Q7),
//     division skipped where the window sum is exactly 0.
//   * order-5 low-pass (scipy.signal.lfilter, state carried across hops), clip, int16 truncation (GriffinLim.py:169-174).
//
// k_gl_blocks: one warp per block; waveform, FFT work buffers and spectra live in shared memory / registers; FP64
//              throughout (the branch cut of angle() makes the iteration chaotic under fp32 round-off).
// k_gl_ola:    gathers the <= 4 blocks covering each output sample in arrival order.
// k_lp_*:      the IIR low-pass as an exact chunked scan over the 5-state recurrence (zero-state chunk pass,
//              sequential 5x5 carry, final pass with clip + int16).
#include <math.h>
#include "common.cuh"
#include "fft.cuh"

namespace sgs {

constexpr int kFft = 256, kHalf = 128, kHop = 160, kBlk = 480, kBins = 129;
constexpr int kGlWarps = 4;

struct GlNodeTables {               // device pointers, built once per node configuration
    const double* window;           // blackman(256)
    const cplx* tw_half;            // exp(-2 pi i t / 128), t < 128
    const cplx* tw_full;            // exp(-2 pi i k / 256), k <= 128
    const int* inv_idx;             // [129][2] mel index of each inverse-mel tap
    const double* inv_w;            // [129][2] weight (0 where unused)
};

struct GlWarpSmem {
    double x[kBlk];
    cplx a[kHalf];
    cplx b[kHalf];
    double zr[2][kBins + 1];
};

__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long item, unsigned idx) {
    // counter-based generator for throughput runs (parity runs pass the reference's MT19937 draws explicitly)
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (item * 480ULL + idx + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(kGlWarps * 32)
k_gl_blocks(const double* __restrict__ logmel, const double* __restrict__ noise, unsigned long long seed,
            double* __restrict__ blocks, const GlNodeTables tab, int n_frames, int n_mels, int first_frame, int iters,
            long long n_items, long long ring_base, int ring_len) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_window = reinterpret_cast<double*>(smem_raw);                 // [256]
    cplx* s_tw_half = reinterpret_cast<cplx*>(s_window + kFft);             // [128]
    cplx* s_tw_full = s_tw_half + kHalf;                                    // [129] (+1 pad)
    GlWarpSmem* ws_all = reinterpret_cast<GlWarpSmem*>(s_tw_full + kBins + 1);
    for (int i = threadIdx.x; i < kFft; i += blockDim.x) s_window[i] = tab.window[i];
    for (int i = threadIdx.x; i < kHalf; i += blockDim.x) s_tw_half[i] = tab.tw_half[i];
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) s_tw_full[i] = tab.tw_full[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GlWarpSmem& ws = ws_all[warp];
    const int per_sess = n_frames - first_frame;                            // blocks per session
    for (long long item = (long long)blockIdx.x * kGlWarps + warp; item < n_items; item += (long long)gridDim.x * kGlWarps) {
        const int sess = (int)(item / per_sess);
        const int k = first_frame + (int)(item - (long long)sess * per_sess);
        const long long frame = (long long)sess * n_frames + k;

        // magnitudes S[f][bin] of the two spectral frames k-1, k for the bins this lane owns: lane + 32 q (q < 4), lane 0 also 128
        double S[2][5];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const double* lm = logmel + (frame - 1 + f) * n_mels;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int bin = (q < 4) ? lane + 32 * q : kHalf;
                double v = 0.0;
                if (q < 4 || lane == 0) {
                    const double w0 = tab.inv_w[bin * 2], w1 = tab.inv_w[bin * 2 + 1];
                    if (w0 != 0.0) v = exp(lm[tab.inv_idx[bin * 2]]) * w0;
                    if (w1 != 0.0) v = fma(exp(lm[tab.inv_idx[bin * 2 + 1]]), w1, v);
                    if (!isfinite(v)) v = 0.0;                              // MelFilterBank.makeNormal
                }
                S[f][q] = v;
            }
        }
        // start waveform
        for (int i = lane; i < kBlk; i += 32)
            ws.x[i] = noise ? noise[frame * kBlk + i] : uniform01(seed, (unsigned long long)(ring_base + frame), (unsigned)i);
        __syncwarp();

        for (int it = 0; it < iters; ++it) {
            // ---- analysis: both frames read the current x ------------------------------------------------
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const int o = f * kHop;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = lane + 32 * i;
                    ws.a[n] = cplx{ws.x[o + 2 * n] * s_window[2 * n], ws.x[o + 2 * n + 1] * s_window[2 * n + 1]};
                }
                __syncwarp();
                fft128<-1>(ws.a, ws.b, s_tw_half, lane);
                // real-FFT split + magnitude projection: zr = S * exp(angle(X)) (real, quirk Q1)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int kk = lane + 32 * q;
                    double re, im;
                    if (kk == 0) {
                        re = ws.a[0].x + ws.a[0].y;                         // DC: numpy returns imag = +0.0 exactly
                        im = 0.0;
                    } else {
                        const cplx A = ws.a[kk], B = cconj(ws.a[kHalf - kk]);
                        const cplx d = csub(A, B), w = s_tw_full[kk];
                        const cplx t = cmul(w, d);                          // e^{-i th}(A - B)
                        re = 0.5 * (A.x + B.x) + 0.5 * t.y;                 // X = (A+B)/2 - i/2 * t
                        im = 0.5 * (A.y + B.y) - 0.5 * t.x;
                    }
                    ws.zr[f][kk] = S[f][q] * exp(atan2(im, re));
                }
                if (lane == 0) {
                    const double re = ws.a[0].x - ws.a[0].y;                // Nyquist, imag = +0.0
                    ws.zr[f][kHalf] = S[f][4] * exp(atan2(0.0, re));
                }
                __syncwarp();
            }
            // ---- synthesis: overwrite x -------------------------------------------------------------------
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const int o = f * kHop;
                // i * e^{+i th} * d = i (cos th + i sin th) d = (-sin th) d + i (cos th) d, with w = (cos th, -sin th):
                // real part = w.y * d, imag part = w.x * d; the spectra are real, so conj() is the identity
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kk = lane + 32 * i;
                    const double A = ws.zr[f][kk], B = ws.zr[f][kHalf - kk];
                    const cplx w = s_tw_full[kk];
                    const double d = A - B;
                    ws.a[kk] = cplx{fma(w.y, d, A + B), w.x * d};
                }
                __syncwarp();
                fft128<+1>(ws.a, ws.b, s_tw_half, lane);
                constexpr double scale = 1.0 / kFft;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = lane + 32 * i;
                    const double r0 = (ws.a[n].x * scale) * s_window[2 * n];
                    const double r1 = (ws.a[n].y * scale) * s_window[2 * n + 1];
                    const int m0 = 2 * n, m1 = 2 * n + 1;
                    if (f == 0) {
                        ws.x[m0] = r0;
                        ws.x[m1] = r1;
                    } else {
                        ws.x[kHop + m0] = (m0 < kFft - kHop) ? ws.x[kHop + m0] + r0 : r0;
                        ws.x[kHop + m1] = (m1 < kFft - kHop) ? ws.x[kHop + m1] + r1 : r1;
                    }
                }
                __syncwarp();
            }
            for (int i = kHop + kFft + lane; i < kBlk; i += 32) ws.x[i] = 0.0;      // istft never reaches [416, 480)
            __syncwarp();
        }
        // batch: one row per (session, frame); streaming: slot of the running frame number in a power-of-two ring
        const long long row = ring_len ? ((ring_base + k) & (ring_len - 1)) : frame;
        for (int i = lane; i < kBlk; i += 32) blocks[row * kBlk + i] = ws.x[i];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// overlap-add + window-sum normalisation, linear-time restatement of the node's ring buffers.
// pos[k] = write head after frame k (host table, the reference's float expression); frame k returns the
// pos[k]-pos[k-1] samples starting at absolute position pos[k] - 480.
// ------------------------------------------------------------------------------------------------
__global__ void k_gl_ola(const double* __restrict__ blocks, const int* __restrict__ pos, const double* __restrict__ ola_window,
                         double* __restrict__ v, int n_frames, int first_frame, long long out_per_sess) {
    const int k = first_frame + blockIdx.x, sess = blockIdx.y;
    const int pk = pos[k], prev = k > 0 ? pos[k - 1] : 0;
    const int shifted = pk - prev;
    const int base_out = prev - (first_frame > 0 ? pos[first_frame - 1] : 0);
    for (int i = threadIdx.x; i < shifted; i += blockDim.x) {
        const int p = pk - kBlk + i;                            // absolute position of this output sample
        double num = 0.0, den = 0.0;
        for (int jj = (k - 4 > first_frame ? k - 4 : first_frame); jj <= k; ++jj) {   // arrival order = ascending frame
            const int off = p - (pos[jj] - kBlk);
            if (off >= 0 && off < kBlk) {
                num += blocks[((long long)sess * n_frames + jj) * kBlk + off];
                den += ola_window[off];
            }
        }
        v[(long long)sess * out_per_sess + base_out + i] = (den != 0.0) ? num / den : num;
    }
}

// ------------------------------------------------------------------------------------------------
// order-ORD IIR (direct form II transposed, scipy.signal.lfilter) along each session's output stream
// ------------------------------------------------------------------------------------------------
constexpr int kLpMaxOrd = 8;
struct LpCoefs { double b[kLpMaxOrd + 1], a[kLpMaxOrd + 1]; int ord; };

__device__ __forceinline__ double lp_step(double xin, double (&z)[kLpMaxOrd], const LpCoefs& c) {
    const double y = fma(c.b[0], xin, z[0]);
#pragma unroll
    for (int i = 0; i < kLpMaxOrd - 1; ++i)
        if (i < c.ord - 1) z[i] = fma(-c.a[i + 1], y, fma(c.b[i + 1], xin, z[i + 1]));
#pragma unroll
    for (int i = 0; i < kLpMaxOrd; ++i)
        if (i == c.ord - 1) z[i] = fma(-c.a[i + 1], y, c.b[i + 1] * xin);
    return y;
}

// pass 1: zero-state end state of each chunk.  thread = (session, chunk)
__global__ void k_lp_state(const double* __restrict__ v, double* __restrict__ states, const __grid_constant__ LpCoefs c,
                           long long n_out, int chunk, int n_chunks, int n_sessions) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)n_chunks * n_sessions) return;
    const int sess = (int)(id / n_chunks), ci = (int)(id - (long long)sess * n_chunks);
    const long long t0 = (long long)ci * chunk, t1 = (t0 + chunk < n_out) ? t0 + chunk : n_out;
    double z[kLpMaxOrd];
#pragma unroll
    for (int i = 0; i < kLpMaxOrd; ++i) z[i] = 0.0;
    const double* p = v + (long long)sess * n_out;
    for (long long t = t0; t < t1; ++t) lp_step(p[t], z, c);
    double* o = states + ((long long)sess * n_chunks + ci) * kLpMaxOrd;
#pragma unroll
    for (int i = 0; i < kLpMaxOrd; ++i) o[i] = z[i];
}

// carry: in place, states[ci] becomes the TRUE state at the START of chunk ci (chunk 0 starts from zi).
// thread = session; phi = A^chunk (ord x ord, row-major).
__global__ void k_lp_carry(double* __restrict__ states, const double* __restrict__ phi, double* __restrict__ zi,
                           int ord, int n_chunks, int n_sessions) {
    const int sess = blockIdx.x * blockDim.x + threadIdx.x;
    if (sess >= n_sessions) return;
    double s[kLpMaxOrd];
    for (int i = 0; i < kLpMaxOrd; ++i) s[i] = (i < ord) ? zi[sess * ord + i] : 0.0;
    double* st = states + (long long)sess * n_chunks * kLpMaxOrd;
    for (int ci = 0; ci < n_chunks; ++ci) {
        double e[kLpMaxOrd], nx[kLpMaxOrd];
        for (int i = 0; i < kLpMaxOrd; ++i) { e[i] = st[ci * kLpMaxOrd + i]; st[ci * kLpMaxOrd + i] = s[i]; }
        for (int i = 0; i < kLpMaxOrd; ++i) {
            double acc = e[i];
            for (int k = 0; k < ord; ++k) acc = fma(phi[i * ord + k], s[k], acc);
            nx[i] = (i < ord) ? acc : 0.0;
        }
        for (int i = 0; i < kLpMaxOrd; ++i) s[i] = nx[i];
    }
    for (int i = 0; i < ord; ++i) zi[sess * ord + i] = s[i];      // final state back to the caller (streaming)
}

// pass 2: filter each chunk from its true state; clip, scale, truncate to int16.
__global__ void k_lp_apply(const double* __restrict__ v, const double* __restrict__ states, short* __restrict__ pcm,
                           double* __restrict__ filtered, const __grid_constant__ LpCoefs c, double norm_div,
                           long long n_out, int chunk, int n_chunks, int n_sessions) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)n_chunks * n_sessions) return;
    const int sess = (int)(id / n_chunks), ci = (int)(id - (long long)sess * n_chunks);
    const long long t0 = (long long)ci * chunk, t1 = (t0 + chunk < n_out) ? t0 + chunk : n_out;
    double z[kLpMaxOrd];
    const double* si = states + ((long long)sess * n_chunks + ci) * kLpMaxOrd;
#pragma unroll
    for (int i = 0; i < kLpMaxOrd; ++i) z[i] = si[i];
    const double* p = v + (long long)sess * n_out;
    for (long long t = t0; t < t1; ++t) {
        const double y = lp_step(p[t], z, c);
        if (filtered) filtered[(long long)sess * n_out + t] = y;
        double q = y / norm_div;                                // np.clip(y / (normFactor * 1.01), -0.99, 0.99) * 32767
        q = q < -0.99 ? -0.99 : (q > 0.99 ? 0.99 : q);
        pcm[(long long)sess * n_out + t] = (short)(int)(q * 32767.0);       // np.int16(): truncation toward zero
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
int gl_blocks_run(const double* logmel, const double* noise, unsigned long long seed, double* blocks, const GlNodeTables& tab,
                  int n_sessions, int n_frames, int n_mels, int first_frame, int iters, long long ring_base, int ring_len,
                  cudaStream_t st) {
    const long long n_items = (long long)n_sessions * (n_frames - first_frame);
    if (n_items <= 0) return SGS_OK;
    const size_t smem = sizeof(double) * kFft + sizeof(cplx) * (kHalf + kBins + 1) + sizeof(GlWarpSmem) * kGlWarps;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_gl_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    long long want = (n_items + kGlWarps - 1) / kGlWarps;
    const int grid = (int)(want < 148 * 5 * 8 ? want : 148 * 5 * 8);
    { ProfScope ps(kProfGlBlocks, st); k_gl_blocks<<<grid, kGlWarps * 32, smem, st>>>(logmel, noise, seed, blocks, tab, n_frames, n_mels, first_frame, iters, n_items, ring_base, ring_len); }
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

int gl_emit_run(const double* blocks, const int* pos, const double* ola_window, double* v, double* states, double* zi,
                const double* phi, const LpCoefs& c, double norm_div, short* pcm, double* filtered, int n_sessions, int n_frames,
                int first_frame, long long n_out, int chunk, int n_chunks, cudaStream_t st) {
    if (n_out <= 0 || n_frames <= first_frame) return SGS_OK;
    { ProfScope ps(kProfGlOla, st); k_gl_ola<<<dim3(n_frames - first_frame, n_sessions), 192, 0, st>>>(blocks, pos, ola_window, v, n_frames, first_frame, n_out); }
    SGS_LAUNCHED();
    ProfScope ps_lp(kProfLowpass, st);
    const long long n_thr = (long long)n_chunks * n_sessions;
    k_lp_state<<<ceil_div(n_thr, 128), 128, 0, st>>>(v, states, c, n_out, chunk, n_chunks, n_sessions);
    SGS_LAUNCHED();
    k_lp_carry<<<ceil_div(n_sessions, 32), 32, 0, st>>>(states, phi, zi, c.ord, n_chunks, n_sessions);
    SGS_LAUNCHED();
    k_lp_apply<<<ceil_div(n_thr, 128), 128, 0, st>>>(v, states, pcm, filtered, c, norm_div, n_out, chunk, n_chunks, n_sessions);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
