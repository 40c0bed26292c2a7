// Batch Griffin-Lim (local/offline.py:131-192, closed form R4 of SURVEY.md 8a') and the audio -> log-mel
// spectrogram (local/offline.py:219-241).
//
// griffin_lim semantics restated: win 800 (periodic Hann = hanning(801)[:-1]), hop 160, 401 bins;
//   x = noise; repeat `iters` times {
//       X_n = rfft(w * x[160n : 160n+800])          for the frames the inverse uses, n < T-5
//       Z_n = S_n * exp(1j * angle(X_n))             S = fromLogMels(spectrogram) through the 401-bin inverse mel
//       re  = zeros(160 T);  re[160n : 160n+800] += irfft(Z_n) * w   for n < T-5 (range(0, 160T-800, 160), quirk Q4)
//       x[:160 T] = re }
//   out = int16(x[:160T] / max|x| * 32767)
// The reference also transforms 5x more STFT frames than it uses and the last 5 spectral frames never reach the
// output; neither is reproduced as work, both are reproduced as results.
//
// One CTA per utterance, 13 warps, one frame per warp per round.  Frames are processed in ascending order, which
// makes the update in place: hop segment h of x is final once frame h is done, and no later frame of the same
// iteration reads it.  The windowed inverse transforms of the last 18 frames sit in a shared-memory ring so each
// output sample is summed over its (up to 5) contributing frames in the reference's accumulation order; a frame's
// ring slot is free until its own result lands there, so it doubles as the second buffer of the Stockham stages.
#include <math.h>
#include "kernels.cuh"

namespace sgs {

constexpr int kBW = 13;                 // warps per CTA (195 frames reach the output at T = 200: 15 full rounds)
constexpr int kN = 800, kM = 400, kHopB = 160, kBinsB = 401, kOverlap = 5;
constexpr int kRingB = kBW + kOverlap;  // 18 slots

// X / |X| without sqrt and divisions: reciprocal square root seed and two Newton steps (<= 2 ulp);
// 0 -> (1, 0) as angle(0) = 0.  |X|^2 outside [2^-900, 2^900] takes the plain route.
__device__ __forceinline__ cplx unit_phase(cplx X) {
    const double m2 = fma(X.x, X.x, X.y * X.y);
    if (__builtin_expect(!(m2 > 0x1p-900 && m2 < 0x1p+900), 0)) {
        const double mag = hypot(X.x, X.y);
        if (mag > 0.0 && isfinite(mag)) return cplx{X.x / mag, X.y / mag};
        return cplx{1.0, 0.0};
    }
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(m2));
    double h = 0.5 * r, e = fma(-m2 * r, h, 0.5);                           // 0.5 - m2 r^2 / 2
    r = fma(r, e, r);
    h = 0.5 * r; e = fma(-m2 * r, h, 0.5);
    r = fma(r, e, r);
    return cplx{X.x * r, X.y * r};
}

__global__ void __launch_bounds__(kBW * 32)
k_gl_batch(const double* __restrict__ logmel /*[B][T][n_mels]*/, double* __restrict__ x /*[B][x_len] in: noise, out: waveform*/,
           const GlBatchTables tab, int T, int n_mels, int iters, long long x_len) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_window = reinterpret_cast<double*>(smem_raw);                 // [800]
    cplx* s_tw_half = reinterpret_cast<cplx*>(s_window + kN);               // [400]
    cplx* s_tw_full = s_tw_half + kM;                                       // [401] (+1 pad)
    double* s_ring = reinterpret_cast<double*>(s_tw_full + kBinsB + 1);     // [18][800]
    cplx* s_work = reinterpret_cast<cplx*>(s_ring + kRingB * kN);           // per warp: a[400]
    double* s_exp = reinterpret_cast<double*>(s_work + kBW * kM);           // per warp: exp(logmel) [n_mels <= 64]
    for (int i = threadIdx.x; i < kN; i += blockDim.x) s_window[i] = tab.window[i];
    fft400_stage_tables(s_tw_half, tab.tw_half);                             // [395] per-stage twiddles
    for (int i = threadIdx.x; i < kBinsB; i += blockDim.x) s_tw_full[i] = tab.tw_full[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cplx* a = s_work + (size_t)warp * kM;
    double* ex = s_exp + warp * 64;
    double* xu = x + (long long)blockIdx.x * x_len;
    const double* lm_u = logmel + (long long)blockIdx.x * T * n_mels;
    const int n_used = T - kOverlap;                                        // frames that reach the output
    constexpr double scale = 1.0 / kN;

    for (int it = 0; it < iters; ++it) {
        for (int g0 = 0; g0 < T; g0 += kBW) {                               // rounds of kBW frames, ascending
            const int n = g0 + warp;
            if (n < n_used) {
                double* slot = s_ring + (size_t)(n % kRingB) * kN;
                cplx* b = reinterpret_cast<cplx*>(slot);                    // scratch until the frame's result is stored
                for (int m = lane; m < n_mels; m += 32) ex[m] = exp(lm_u[(long long)n * n_mels + m]);
                const double* xin = xu + (long long)n * kHopB;
                for (int i = lane; i < kM; i += 32) {
                    const double2 xv = *reinterpret_cast<const double2*>(xin + 2 * i);
                    const double2 wv = *reinterpret_cast<const double2*>(s_window + 2 * i);
                    a[i] = cplx{xv.x * wv.x, xv.y * wv.y};
                }
                __syncwarp();
                fft400<-1>(a, b, s_tw_half, lane);
                // real-FFT split, phase projection Z = S * X/|X|, and the inverse split, bin pairs (k, 400-k) together
                for (int k = lane; k <= kM / 2; k += 32) {
                    const int k2 = kM - k;
                    cplx Xk, Xk2;
                    if (k == 0) {
                        Xk = cplx{a[0].x + a[0].y, 0.0};                    // DC
                        Xk2 = cplx{a[0].x - a[0].y, 0.0};                   // Nyquist (k2 = 400)
                    } else {
                        const cplx A = a[k], B = cconj(a[k2]);
                        const cplx d = csub(A, B);
                        const cplx t1 = cmul(s_tw_full[k], d);
                        Xk = cplx{0.5 * (A.x + B.x) + 0.5 * t1.y, 0.5 * (A.y + B.y) - 0.5 * t1.x};
                        // X[400-k] from the same pair: A' = a[k2], B' = conj(a[k])
                        const cplx A2 = a[k2], B2 = cconj(a[k]);
                        const cplx d2 = csub(A2, B2);
                        const cplx t2 = cmul(s_tw_full[k2], d2);
                        Xk2 = cplx{0.5 * (A2.x + B2.x) + 0.5 * t2.y, 0.5 * (A2.y + B2.y) - 0.5 * t2.x};
                    }
                    auto project = [&](cplx X, int bin) -> cplx {
                        const double w0 = tab.inv_w[bin * 2], w1 = tab.inv_w[bin * 2 + 1];
                        double S = 0.0;
                        if (w0 != 0.0) S = ex[tab.inv_idx[bin * 2]] * w0;
                        if (w1 != 0.0) S = fma(ex[tab.inv_idx[bin * 2 + 1]], w1, S);
                        if (!isfinite(S)) S = 0.0;
                        const cplx u = unit_phase(X);
                        return cplx{S * u.x, S * u.y};
                    };
                    const cplx Zk = project(Xk, k), Zk2 = project(Xk2, k2);
                    // irfft ignores the imaginary parts of the DC and Nyquist bins
                    if (k == 0) {
                        a[0] = cplx{Zk.x + Zk2.x, Zk.x - Zk2.x};
                    } else {
                        // Zin[k] = (A + B) + i e^{+i th_k} (A - B) with A = Z[k], B = conj(Z[400-k]); same for 400-k
                        const cplx A = Zk, B = cconj(Zk2);
                        const cplx s1 = cadd(A, B), d1 = csub(A, B);
                        const cplx w = s_tw_full[k];                        // (cos, -sin)
                        // i e^{i th} d = i (cos + i sin)(dx + i dy) = (-sin dx - cos dy) + i (cos dx - sin dy)
                        const cplx r1 = cplx{fma(w.y, d1.x, -w.x * d1.y), fma(w.x, d1.x, w.y * d1.y)};
                        const cplx A2 = Zk2, B2 = cconj(Zk);
                        const cplx s2 = cadd(A2, B2), d2 = csub(A2, B2);
                        const cplx w2 = s_tw_full[k2];
                        const cplx r2 = cplx{fma(w2.y, d2.x, -w2.x * d2.y), fma(w2.x, d2.x, w2.y * d2.y)};
                        a[k] = cadd(s1, r1);
                        if (k2 != k) a[k2] = cadd(s2, r2);
                    }
                }
                __syncwarp();
                fft400<+1>(a, b, s_tw_half, lane);
                for (int i = lane; i < kM; i += 32) {
                    const double2 wv = *reinterpret_cast<const double2*>(s_window + 2 * i);
                    const cplx v = a[i];
                    *reinterpret_cast<double2*>(slot + 2 * i) = make_double2((v.x * scale) * wv.x, (v.y * scale) * wv.y);
                }
            }
            __syncthreads();
            // hop segments g0 .. g0+7 are final now: sum their contributing frames in ascending order
            for (int idx = threadIdx.x; idx < kBW * kHopB; idx += blockDim.x) {
                const int h = g0 + idx / kHopB, j = idx - (idx / kHopB) * kHopB;
                if (h >= T) break;
                double acc = 0.0;
                for (int nn = h - (kOverlap - 1); nn <= h; ++nn)
                    if (nn >= 0 && nn < n_used) acc += s_ring[(size_t)(nn % kRingB) * kN + (h - nn) * kHopB + j];
                xu[(long long)h * kHopB + j] = acc;
            }
            __syncthreads();
        }
    }
}

// max |x| over the first n samples of each utterance, then int16(x / max * 32767)
__global__ void k_absmax(const double* __restrict__ x, long long x_len, long long n, double* __restrict__ out) {
    __shared__ double red[32];
    const double* xu = x + (long long)blockIdx.x * x_len;
    double m = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(xu[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[blockIdx.x] = m;
    }
}

__global__ void k_scale_int16(const double* __restrict__ x, long long x_len, long long n, const double* __restrict__ mx,
                              short* __restrict__ pcm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int u = blockIdx.y;
    pcm[(long long)u * n + i] = (short)(int)((x[(long long)u * x_len + i] / mx[u]) * 32767.0);
}

int gl_batch_run(const double* logmel, double* x, const GlBatchTables& tab, int n_utt, int T, int n_mels, int iters,
                 long long x_len, double* mx, short* pcm, cudaStream_t st) {
    const size_t smem = sizeof(double) * kN + sizeof(cplx) * (kM + kBinsB + 1) + sizeof(double) * kRingB * kN +
                        sizeof(cplx) * kBW * kM + sizeof(double) * kBW * 64;
    static unsigned long long optin = 0;
    SGS_CUDA(smem_optin(k_gl_batch, smem, &optin));
    {
        ProfScope ps(kProfGlBatch, st);
        k_gl_batch<<<n_utt, kBW * 32, smem, st>>>(logmel, x, tab, T, n_mels, iters, x_len);
    }
    SGS_LAUNCHED();
    const long long n = (long long)T * kHopB;
    k_absmax<<<n_utt, 256, 0, st>>>(x, x_len, n, mx);
    SGS_LAUNCHED();
    k_scale_int16<<<dim3(ceil_div(n, 256), n_utt), 256, 0, st>>>(x, x_len, n, mx, pcm);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// ------------------------------------------------------------------------------------------------
// audio -> log-mel (compute_spectrogram with 16 ms windows as train.py:128 calls it: window 256, shift 160):
// frame i = hann256 * audio[160 i - 96 : 160 i + 160] (96 zeros in front), |rfft|, log(mel . |X| + 1e-7)
// one warp per frame
// ------------------------------------------------------------------------------------------------
constexpr int kSpecWarps = 4;
__global__ void __launch_bounds__(kSpecWarps * 32)
k_logmel(const double* __restrict__ audio, long long n_audio, const double* __restrict__ window /*[256]*/,
         const cplx* __restrict__ tw_half /*128*/, const cplx* __restrict__ tw_full /*129*/,
         const double* __restrict__ mel /*[129][n_mels]*/, int n_mels, long long n_frames, int shift, int pad,
         double* __restrict__ out /*[n_frames][n_mels]*/) {
    __shared__ cplx s_a[kSpecWarps][128], s_b[kSpecWarps][128], s_tw[128];
    __shared__ double s_mag[kSpecWarps][130];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_tw[i] = tw_half[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long f = (long long)blockIdx.x * kSpecWarps + warp; f < n_frames; f += (long long)gridDim.x * kSpecWarps) {
        const long long t0 = f * shift - pad;
        for (int i = lane; i < 128; i += 32) {
            const long long t = t0 + 2 * i;
            const double v0 = (t >= 0 && t < n_audio) ? audio[t] : 0.0;
            const double v1 = (t + 1 >= 0 && t + 1 < n_audio) ? audio[t + 1] : 0.0;
            s_a[warp][i] = cplx{v0 * window[2 * i], v1 * window[2 * i + 1]};
        }
        __syncwarp();
        fft128<-1>(s_a[warp], s_b[warp], s_tw, lane);
        for (int k = lane; k <= 128; k += 32) {
            double re, im;
            if (k == 0) { re = s_a[warp][0].x + s_a[warp][0].y; im = 0.0; }
            else if (k == 128) { re = s_a[warp][0].x - s_a[warp][0].y; im = 0.0; }
            else {
                const cplx A = s_a[warp][k], B = cconj(s_a[warp][128 - k]);
                const cplx t = cmul(tw_full[k], csub(A, B));
                re = 0.5 * (A.x + B.x) + 0.5 * t.y;
                im = 0.5 * (A.y + B.y) - 0.5 * t.x;
            }
            s_mag[warp][k] = hypot(re, im);
        }
        __syncwarp();
        for (int m = lane; m < n_mels; m += 32) {
            double acc = 0.0;
            for (int k = 0; k <= 128; ++k) acc = fma(s_mag[warp][k], mel[k * n_mels + m], acc);
            double v = log(acc + 0.0000001);
            if (!isfinite(v)) v = 0.0;
            out[f * n_mels + m] = v;
        }
        __syncwarp();
    }
}

int logmel_run(const double* audio, long long n_audio, const double* window, const cplx* tw_half, const cplx* tw_full,
               const double* mel, int n_mels, long long n_frames, int shift, int pad, double* out, cudaStream_t st) {
    if (n_frames <= 0) return SGS_OK;
    long long want = (n_frames + kSpecWarps - 1) / kSpecWarps;
    const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
    {
        ProfScope ps(kProfLogMel, st);
        k_logmel<<<grid, kSpecWarps * 32, 0, st>>>(audio, n_audio, window, tw_half, tw_full, mel, n_mels, n_frames, shift, pad, out);
    }
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
