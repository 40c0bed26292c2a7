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
// One launch per iteration (the waveform ping-pongs between two buffers), three CTAs per SM, each working through an equal
// run of 8-frame rounds of the utterance-major line of all rounds - see k_gl_batch_iter.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "kernels.cuh"

namespace sgs {

constexpr int kN = 800, kM = 400, kHopB = 160, kBinsB = 401, kOverlap = 5;
constexpr int kBF = 8;                  // frames per round
constexpr int kBT = kBF * 20;           // threads per CTA: one 20-point transform per thread and pass
constexpr int kRow = 21;                // row stride of a frame buffer (20 complex entries + 1 pad: rows and columns conflict-free)
constexpr int kBufC = 20 * kRow;        // complex entries per frame buffer
constexpr int kPairsPerThread = kBF * 200 / kBT;   // bin pairs (k, 400 - k), k = 1..200, of the round's frames per thread
constexpr int kCarry = kOverlap - 1;    // hop segments that still receive contributions from later rounds

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

// 20-point DFT in registers by the prime-factor map (4 x 5, no twiddles between the two passes):
//   g[a][b] = v[(5 a + 4 b) % 20];  5-point DFTs along b, 4-point DFTs along a;  X[(5 ka + 16 kb) % 20] = g[ka][kb]
// so X[k] ends in v[pos20(k)], pos20(k) = (5 (k % 4) + 4 (k % 5)) % 20.  224 fp64 operations.
__device__ __forceinline__ constexpr int pos20(int k) { return (5 * (k % 4) + 4 * (k % 5)) % 20; }
template <int SIGN>
__device__ __forceinline__ void dft20(cplx (&v)[20]) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        cplx t[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) t[b] = v[(5 * a + 4 * b) % 20];
        Butterfly<5, SIGN>::run(t);
#pragma unroll
        for (int b = 0; b < 5; ++b) v[(5 * a + 4 * b) % 20] = t[b];
    }
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        cplx t[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) t[a] = v[(5 * a + 4 * b) % 20];
        Butterfly<4, SIGN>::run(t);
#pragma unroll
        for (int a = 0; a < 4; ++a) v[(5 * a + 4 * b) % 20] = t[a];
    }
}

// magnitudes S = fromLogMels(spectrogram) of the frames the inverse uses: [utterance][frame < T-5][401]
__global__ void k_gl_batch_mag(const double* __restrict__ logmel, const int* __restrict__ inv_idx, const double* __restrict__ inv_w,
                               double* __restrict__ S, int T, int n_used, int n_mels) {
    __shared__ double ex[64];
    const int n = blockIdx.x, u = blockIdx.y;
    const double* lm = logmel + ((long long)u * T + n) * n_mels;
    for (int m = threadIdx.x; m < n_mels; m += blockDim.x) ex[m] = exp(lm[m]);
    __syncthreads();
    double* out = S + ((long long)u * n_used + n) * kBinsB;
    for (int b = threadIdx.x; b < kBinsB; b += blockDim.x) {
        const double w0 = inv_w[b * 2], w1 = inv_w[b * 2 + 1];
        double v = 0.0;
        if (w0 != 0.0) v = ex[inv_idx[b * 2]] * w0;
        if (w1 != 0.0) v = fma(ex[inv_idx[b * 2 + 1]], w1, v);
        out[b] = isfinite(v) ? v : 0.0;                                     // MelFilterBank.makeNormal
    }
}

// One Griffin-Lim iteration: xin -> xout.
//
// The 800-point real transform is a 400-point complex one laid out as 20 x 20, and a thread holds one 20-point transform in
// registers: pass 1 (thread = (frame, l)) transforms z[l + 20 m] over m and applies W400^(l k1), pass 2 (thread = (frame,
// k1)) transforms over l; bin k1 + 20 k2 then sits at [k1][k2] of the frame's shared-memory buffer, which every pass
// updates in place (a thread writes exactly the 20 entries it read).  The real-FFT split, the phase projection
// Z = S X / |X| and the inverse split run per bin pair (k, 400 - k); the inverse mirrors the two passes.  A round handles
// kBF consecutive frames; their windowed outputs are summed per hop segment in ascending frame order - the order of the
// reference's `re[i:i+800] += ...` loop - on top of the partial sums the earlier rounds left for the 4 segments that were
// still open (`carry`), so no frame buffer outlives its round.  The first version ran four Stockham stages per transform
// through shared memory with 13 frames in flight per SM: 22 % of the FP64 pipe (profiles/ncu_gl_blocks8_r01b.txt).
//
// Work split: the rounds of all utterances form one line (utterance-major, `rpu` rounds per utterance) cut into runs of
// `slots_per_cta`, one run per CTA.  With many utterances a run is exactly one utterance (4096 CTAs for config 4: the
// hardware's dynamic CTA placement evens out the SMs; a perfectly even static split over one wave of CTAs measured 117.0 ms
// against 112.1 ms, and splitting only the last utterances of the launch changed nothing - the few CTAs left at the end
// run proportionally faster).  With fewer utterances than a few waves of CTAs the runs are equal shares of the line, so that
// 512 utterances (config 4 over 8 GPUs) keep every SM busy: 15.4 ms against 16.6 ms in whole or quartered utterances.
// A run that starts inside an utterance, at round r, owns its segments from 8 r + 4 on: it re-computes frames 8 r .. 8 r + 3,
// which its predecessor also computes - in one extra, half-empty round - to finish segments 8 r .. 8 r + 3.  No sum crosses
// a CTA and every sample is summed in the same order wherever the cuts fall, so the output does not depend on the batch.
__global__ void __launch_bounds__(kBT, 3)
k_gl_batch_iter(const double* __restrict__ xin, long long in_stride, double* __restrict__ xout, long long out_stride,
                const double* __restrict__ S, const GlBatchTables tab, int T, int n_used, int rpu, int slots_per_cta,
                long long total_slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* s_buf = reinterpret_cast<cplx*>(smem_raw);                        // [kBF][kBufC]
    double* s_carry = reinterpret_cast<double*>(s_buf + kBF * kBufC);       // [kCarry][160]
    cplx* s_twt = reinterpret_cast<cplx*>(s_carry + kCarry * kHopB);        // [20][20] W400^(l k1) (symmetric)
    double2* s_win2 = reinterpret_cast<double2*>(s_twt + 400);              // [400] window pairs (w[2n], w[2n+1])
    cplx* s_twf = reinterpret_cast<cplx*>(s_win2 + kM);                     // [201] exp(-2 pi i k / 800)
    const int tid = threadIdx.x;
    for (int i = tid; i < 400; i += kBT) s_twt[i] = tab.tw_half[((i / 20) * (i % 20)) % kM];
    for (int i = tid; i < kM; i += kBT) s_win2[i] = make_double2(tab.window[2 * i], tab.window[2 * i + 1]);
    for (int i = tid; i <= kM / 2; i += kBT) s_twf[i] = tab.tw_full[i];
    __syncthreads();

    const int f = tid / 20, l = tid - f * 20;                               // frame of the round, row / column index
    cplx* A = s_buf + f * kBufC;
    constexpr double scale = 1.0 / kN;
    long long q = (long long)blockIdx.x * slots_per_cta;
    const long long q_end = min(q + slots_per_cta, total_slots);
    while (q < q_end) {
    const int u = (int)(q / rpu), r0 = (int)(q - (long long)u * rpu);
    const bool to_end = q_end - q >= rpu - r0;                               // the run covers the rest of this utterance
    const int h0 = r0 ? r0 * kBF + kCarry : 0;
    const int h1 = to_end ? T : min(T, (r0 + (int)(q_end - q)) * kBF + kCarry);
    q += to_end ? rpu - r0 : q_end - q;
    const double* xu = xin + (long long)u * in_stride;
    double* xo = xout + (long long)u * out_stride;
    const double* Su = S + (long long)u * n_used * kBinsB;
    const int n_lim = min(n_used, h1);                                      // frames >= h1 only reach segments of the next run
#pragma unroll
    for (int i = 0; i < kCarry; ++i) s_carry[i * kHopB + tid] = 0.0;        // column tid of the carry belongs to thread tid

    // the raw samples of this thread's frame, requested one round ahead: DRAM latency then overlaps the overlap-add and the
    // barrier instead of opening pass 1
    double2 xr[20];
    if (r0 * kBF + f < n_lim) {
        const double2* x2 = reinterpret_cast<const double2*>(xu + (long long)(r0 * kBF + f) * kHopB) + l;
#pragma unroll
        for (int m = 0; m < 20; ++m) xr[m] = x2[20 * m];
    }
    for (int g0 = r0 * kBF; g0 < h1; g0 += kBF) {
        const int n = g0 + f;
        const bool live = n < n_lim;
        cplx v[20];
        {
            // the next round's magnitudes come from DRAM (an utterance's 0.9 MB per iteration does not stay in L2 across the 440
            // utterances in flight): ask for the lines now
            const int gn = g0 + kBF;
            if (gn < n_lim) {
                const char* ps = reinterpret_cast<const char*>(Su + (long long)gn * kBinsB);
                const int s_bytes = (int)sizeof(double) * kBinsB * min(kBF, n_lim - gn);
                for (int o = 128 * tid; o < s_bytes; o += 128 * kBT) asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + o));
            }
        }
        // ---- pass 1: window + pack z[j] = x[2j] w[2j] + i x[2j+1] w[2j+1], j = l + 20 m; transform over m; twiddle ----------
        if (live) {
#pragma unroll
            for (int m = 0; m < 20; ++m) {
                const double2 wv = s_win2[l + 20 * m];
                v[m] = cplx{xr[m].x * wv.x, xr[m].y * wv.y};
            }
            dft20<-1>(v);
            A[l] = v[pos20(0)];
#pragma unroll
            for (int k1 = 1; k1 < 20; ++k1) A[k1 * kRow + l] = cmul(v[pos20(k1)], s_twt[k1 * 20 + l]);
        }
        __syncthreads();
        // ---- pass 2: row k1 = l, transform over the column index; bin k1 + 20 k2 -> [k1][k2] ---------------------------------
        if (live) {
#pragma unroll
            for (int j = 0; j < 20; ++j) v[j] = A[l * kRow + j];
            dft20<-1>(v);
#pragma unroll
            for (int k2 = 0; k2 < 20; ++k2) A[l * kRow + k2] = v[pos20(k2)];
        }
        // the magnitudes of this thread's 10 bin pairs: requested before the barrier, so DRAM latency overlaps the wait
        double sa[kPairsPerThread], sb[kPairsPerThread];
#pragma unroll
        for (int i = 0; i < kPairsPerThread; ++i) {
            const int p = tid + kBT * i, fp = p / 200, k = 1 + (p - fp * 200), np = g0 + fp;
            const bool ok = np < n_lim;
            sa[i] = ok ? Su[(long long)np * kBinsB + k] : 0.0;
            sb[i] = ok ? Su[(long long)np * kBinsB + kM - k] : 0.0;
        }
        __syncthreads();
        // ---- bin pairs (k, 400 - k), k = 1..200: split, Z = S X / |X|, inverse split -------------------------------------------
        // with A = Zc[k], B = conj(Zc[400-k]), t = w_k (A - B):  X[k] = E + O, X[400-k] = conj(E - O),  E = (A + B)/2, O = -i t/2
        // and back:  Zin[k] = s + r, Zin[400-k] = conj(s - r),  s = Z[k] + conj(Z[400-k]),  r = i conj(w_k) (Z[k] - conj(Z[400-k]))
#pragma unroll
        for (int i = 0; i < kPairsPerThread; ++i) {
            const int p = tid + kBT * i, fp = p / 200, k = 1 + (p - fp * 200), k2 = kM - k;
            if (g0 + fp >= n_lim) continue;
            cplx* Ap = s_buf + fp * kBufC;
            const int ia = (k % 20) * kRow + k / 20, ib = (k2 % 20) * kRow + k2 / 20;
            const double sk = sa[i], sk2 = sb[i];
            const cplx Za = Ap[ia], Zb = Ap[ib], w = s_twf[k];
            const cplx d = cplx{Za.x - Zb.x, Za.y + Zb.y}, t = cmul(w, d);
            // 2 X instead of X: X / |X| does not see the scale, and a factor of two passes through the reciprocal-square-root
            // seed and the Newton steps of unit_phase exactly - the same bits without the halves
            const double ex = Za.x + Zb.x, ey = Za.y - Zb.y;
            const cplx X1 = cplx{ex + t.y, ey - t.x};
            const cplx X2 = cplx{ex - t.y, -ey - t.x};
            const cplx u1 = unit_phase(X1), u2 = unit_phase(X2);
            const cplx Z1 = cplx{sk * u1.x, sk * u1.y}, Z2 = cplx{sk2 * u2.x, sk2 * u2.y};
            const cplx sm = cplx{Z1.x + Z2.x, Z1.y - Z2.y}, df = cplx{Z1.x - Z2.x, Z1.y + Z2.y};
            // i e^{i th} df = i (cos + i sin)(dx + i dy) = (-sin dx - cos dy) + i (cos dx - sin dy), w = (cos, -sin)
            const cplx r = cplx{fma(w.y, df.x, -w.x * df.y), fma(w.x, df.x, w.y * df.y)};
            Ap[ia] = cplx{sm.x + r.x, sm.y + r.y};
            if (k2 != k) Ap[ib] = cplx{sm.x - r.x, -(sm.y - r.y)};
        }
        if (tid < kBF && g0 + tid < n_lim) {
            // DC and Nyquist are real: Z = S sign(X) (the reference's exp(1j * pi) has real part -1; irfft drops the imaginary parts)
            cplx* Ap = s_buf + tid * kBufC;
            const long long so = (long long)(g0 + tid) * kBinsB;
            const cplx z0 = Ap[0];
            const cplx ud = unit_phase(cplx{z0.x + z0.y, 0.0}), un = unit_phase(cplx{z0.x - z0.y, 0.0});
            const double zd = Su[so] * ud.x, zn = Su[so + kM] * un.x;
            Ap[0] = cplx{zd + zn, zd - zn};
        }
        __syncthreads();
        // ---- inverse pass A: row k1 = l over k2, twiddle conj(W400^(l' k1)); inverse pass B: column l over k1 -------------------
        if (live) {
#pragma unroll
            for (int j = 0; j < 20; ++j) v[j] = A[l * kRow + j];
            dft20<+1>(v);
            A[l * kRow] = v[pos20(0)];
#pragma unroll
            for (int j = 1; j < 20; ++j) {
                const cplx w = s_twt[j * 20 + l];                            // = W400^(j k1), the table is symmetric
                A[l * kRow + j] = cmul(v[pos20(j)], cplx{w.x, -w.y});
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int k1 = 0; k1 < 20; ++k1) v[k1] = A[k1 * kRow + l];
            dft20<+1>(v);
            // y[2j], y[2j+1] = Re, Im of the unnormalised inverse at j = l + 20 m, windowed: kept at [m][l]
#pragma unroll
            for (int m = 0; m < 20; ++m) {
                const double2 wv = s_win2[l + 20 * m];
                A[m * kRow + l] = cplx{(v[pos20(m)].x * scale) * wv.x, (v[pos20(m)].y * scale) * wv.y};
            }
        }
        {
            // (assigned on every path, so that the registers are free between pass 1 and here)
            const bool more = n + kBF < n_lim;
            const double2* x2 = reinterpret_cast<const double2*>(xu + (long long)(more ? n + kBF : 0) * kHopB) + l;
#pragma unroll
            for (int m = 0; m < 20; ++m) xr[m] = make_double2(0.0, 0.0);
            if (more) {
#pragma unroll
                for (int m = 0; m < 20; ++m) xr[m] = x2[20 * m];
            }
        }
        __syncthreads();
        // ---- overlap-add: segments g0 .. g0+kBF-1 are final, the next 4 stay open ---------------------------------------------------
        {
            const int j = tid;                                               // sample within the hop (kBT == kHopB)
            const double* Ad = reinterpret_cast<const double*>(s_buf);
#pragma unroll
            for (int sgm = 0; sgm < kBF + kCarry; ++sgm) {
                const int h = g0 + sgm;
                double acc = sgm < kCarry ? s_carry[sgm * kHopB + j] : 0.0;
#pragma unroll
                for (int d = kOverlap - 1; d >= 0; --d) {                    // ascending frame h - d
                    const int fr = sgm - d;
                    if (fr >= 0 && fr < kBF && g0 + fr < n_lim) {
                        const int i = d * kHopB + j, jc = i >> 1;
                        acc += Ad[(size_t)fr * (2 * kBufC) + 2 * ((jc / 20) * kRow + jc % 20) + (i & 1)];
                    }
                }
                if (sgm < kBF) { if (h >= h0 && h < h1) xo[(long long)h * kHopB + j] = acc; }
                else s_carry[(sgm - kBF) * kHopB + j] = acc;
            }
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
    static_assert(kBT == kHopB, "the overlap-add maps one thread to one sample of a hop");
    const long long n = (long long)T * kHopB;
    const int n_used = T - kOverlap;
    if (iters > 0) {
        const size_t smem = sizeof(cplx) * (kBF * kBufC + 400 + kM / 2 + 1) + sizeof(double) * kCarry * kHopB + sizeof(double2) * kM;
        static unsigned long long optin = 0;
        SGS_CUDA(smem_optin(k_gl_batch_iter, smem, &optin));
        int dev = 0, sms = 148;
        SGS_CUDA(cudaGetDevice(&dev));
        SGS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int rpu = (T + kBF - 1) / kBF;                                 // rounds per utterance
        const long long slots = 3LL * sms, total = (long long)n_utt * rpu;   // CTAs resident at once; rounds in all
        int per_cta = rpu;                                                   // whole utterances ...
        if (n_utt < 4 * slots) per_cta = (int)std::max<long long>((total + slots - 1) / slots, std::min(rpu, 2));   // ... or equal runs
        const int grid = (int)((total + per_cta - 1) / per_cta);
        double *y = nullptr, *S = nullptr;
        SGS_CUDA(cudaMallocAsync((void**)&y, sizeof(double) * (size_t)n_utt * n, st));
        SGS_CUDA(cudaMallocAsync((void**)&S, sizeof(double) * (size_t)n_utt * n_used * kBinsB, st));
        k_gl_batch_mag<<<dim3(n_used, n_utt), 128, 0, st>>>(logmel, tab.inv_idx, tab.inv_w, S, T, n_used, n_mels);
        SGS_LAUNCHED();
        {
            ProfScope ps(kProfGlBatch, st);
            for (int it = 0; it < iters; ++it) {
                const bool fwd = (it & 1) == 0;                              // x -> y on even iterations, y -> x on odd ones
                k_gl_batch_iter<<<grid, kBT, smem, st>>>(fwd ? x : y, fwd ? x_len : n, fwd ? y : x, fwd ? n : x_len, S, tab, T, n_used,
                                                         rpu, per_cta, total);
                SGS_LAUNCHED();
            }
        }
        if (iters & 1)
            SGS_CUDA(cudaMemcpy2DAsync(x, sizeof(double) * x_len, y, sizeof(double) * n, sizeof(double) * n, n_utt, cudaMemcpyDeviceToDevice, st));
        SGS_CUDA(cudaFreeAsync(y, st));
        SGS_CUDA(cudaFreeAsync(S, st));
    }
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
