// Stateful streaming kernels behind the livenodes nodes (10 ms-frame latency path).
//
// Reference semantics restated:
//   ECogFeatCalc chain .... livenodes/ECogFeatCalc.py:67-104 over livenodes/FrameBuffer.py:60-177: three causal
//                           sosfilt stages with carried state, 50 ms frames every 10 ms (fractional positions
//                           rounded by the host exactly as FrameBuffer.py:177 does), log(sum(x^2)+0.01), 21-row
//                           stack buffer read at rows [0,5,10,15,20] (zeros before the stream start).
//   Dequantization ........ livenodes/Dequantization.py:15-18.
//   GriffinLimSynthesis ... livenodes/GriffinLim.py:98-174 (ring overlap-add, window-sum normalisation, lfilter
//                           with carried state, int16).
// One small launch per pushed chunk / frame; all state stays resident in device memory between calls.  The
// filters run one thread per channel here: a 64-sample packet is ~15 us of serial work, latency not throughput.
#include <math.h>
#include "kernels.cuh"

namespace sgs {


template <int NB, typename TIn>
__global__ void __launch_bounds__(128)
k_feat_stream(const TIn* __restrict__ x, int n, int n_channels, long long t0 /*samples consumed before this call*/,
              double* __restrict__ z /*[2*NB][C]*/, double* __restrict__ sq_ring /*[kSqRing][C]*/,
              double* __restrict__ feat_ring /*[kFeatRing][C]*/, const double* __restrict__ zero_fill_resp, int zero_fill,
              int cold_last, int frame_size, int order, int step, double* __restrict__ out /*[frames][C*(order+1)]*/,
              const __grid_constant__ FeatCoefs cf, const __grid_constant__ StreamFrames fr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_channels) return;
    constexpr int NF = NB / kSecPerFilter;
    double z0[NB], z1[NB];
    int first = 0;
    if (t0 == 0) {
        // cold start on the very first sample (FrameBuffer.py:87-98): filter f starts from zi * (its first input),
        // the last one from its warm-started state - or cold like the others (ECogFeatCalc(warm_start=False))
        double v = (double)x[c];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
#pragma unroll
            for (int s = 0; s < kSecPerFilter; ++s) {
                const int i = f * kSecPerFilter + s;
                const bool warm = f == NF - 1 && !cold_last;
                z0[i] = warm ? cf.zi_warm[s][0] : cf.zi[i][0] * v;
                z1[i] = warm ? cf.zi_warm[s][1] : cf.zi[i][1] * v;
            }
#pragma unroll
            for (int s = 0; s < kSecPerFilter; ++s) {
                const int i = f * kSecPerFilter + s;
                const double y = fma(cf.c[i][0], v, z0[i]);
                z0[i] = fma(cf.c[i][1], v, fma(-cf.c[i][3], y, z1[i]));
                z1[i] = fma(-cf.c[i][4], y, cf.c[i][2] * v);
                v = y;
            }
        }
        sq_ring[(0 & (kSqRing - 1)) * n_channels + c] = v * v;
        first = 1;
    } else {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            z0[i] = z[(2 * i) * n_channels + c];
            z1[i] = z[(2 * i + 1) * n_channels + c];
        }
    }
    for (int t = first; t < n; ++t) {
        double v = (double)x[(long long)t * n_channels + c];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const double y = fma(cf.c[i][0], v, z0[i]);
            z0[i] = fma(cf.c[i][1], v, fma(-cf.c[i][3], y, z1[i]));
            z1[i] = fma(-cf.c[i][4], y, cf.c[i][2] * v);
            v = y;
        }
        sq_ring[(int)((t0 + t) & (kSqRing - 1)) * n_channels + c] = v * v;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        z[(2 * i) * n_channels + c] = z0[i];
        z[(2 * i + 1) * n_channels + c] = z1[i];
    }
    // frames completed inside this push
    const int width = n_channels * (order + 1);
    for (int q = 0; q < fr.n; ++q) {
        const long long e = fr.end[q], k = fr.index[q];
        double acc = 0.0;
        for (long long t = e - frame_size; t < e; ++t) {
            if (t < 0) { const double r = zero_fill_resp[t + zero_fill]; acc = fma(r, r, acc); }
            else acc += sq_ring[(int)(t & (kSqRing - 1)) * n_channels + c];
        }
        const double f = log(acc + 0.01);
        feat_ring[(int)(k & (kFeatRing - 1)) * n_channels + c] = f;
        for (int tap = 0; tap <= order; ++tap) {
            const long long kk = k - (long long)(order - tap) * step;
            out[(long long)q * width + c * (order + 1) + tap] =
                (kk < 0) ? 0.0 : (kk == k ? f : feat_ring[(int)(kk & (kFeatRing - 1)) * n_channels + c]);
        }
    }
}

int feat_stream_run(int n_biquads, const void* x, bool x_is_f64, int n, int n_channels, long long t0, double* z,
                    double* sq_ring, double* feat_ring, const double* zf, int zero_fill, int cold_last, int frame_size, int order,
                    int step, double* out, const FeatCoefs& cf, const StreamFrames& fr, cudaStream_t st) {
    const int grid = ceil_div(n_channels, 128);
    ProfScope ps(kProfStream, st);
#define SGS_LAUNCH(NB, T) k_feat_stream<NB, T><<<grid, 128, 0, st>>>((const T*)x, n, n_channels, t0, z, sq_ring, feat_ring, zf, \
                                                                     zero_fill, cold_last, frame_size, order, step, out, cf, fr)
    if (n_biquads == 24) { if (x_is_f64) SGS_LAUNCH(24, double); else SGS_LAUNCH(24, float); }
    else if (n_biquads == 16) { if (x_is_f64) SGS_LAUNCH(16, double); else SGS_LAUNCH(16, float); }
    else { set_error("unsupported biquad count %d", n_biquads); return SGS_ERR_UNSUPPORTED; }
#undef SGS_LAUNCH
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// ---- filtering FrameBuffer (livenodes/FrameBuffer.py:86-143): one sosfilt cascade with carried state -------------------------------
// thread = channel; scipy's _sosfilt recurrence with separate multiplies and adds (y = b0 x + z0; z0 = b1 x - a1 y + z1;
// z1 = b2 x - a2 y), sections inner loop, samples outer.  first != 0: the state starts from zi (warm start) or zi * x[0] (cold).
template <typename TIn>
__global__ void __launch_bounds__(128)
k_sos_stream(const TIn* __restrict__ x, int n, int n_channels, double* __restrict__ z /*[2*sections][C]*/, double* __restrict__ y,
             int first, int warm_start, const __grid_constant__ SosCoefs cf) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_channels) return;
    double z0[kSosMaxSections], z1[kSosMaxSections];
    const double x0 = (double)x[c];
#pragma unroll
    for (int s = 0; s < kSosMaxSections; ++s) {
        if (s < cf.n_sections) {
            z0[s] = first ? (warm_start ? cf.zi[s][0] : cf.zi[s][0] * x0) : z[(2 * s) * n_channels + c];
            z1[s] = first ? (warm_start ? cf.zi[s][1] : cf.zi[s][1] * x0) : z[(2 * s + 1) * n_channels + c];
        } else { z0[s] = 0.0; z1[s] = 0.0; }
    }
    for (int t = 0; t < n; ++t) {
        double v = (double)x[(long long)t * n_channels + c];
#pragma unroll
        for (int s = 0; s < kSosMaxSections; ++s) {
            if (s < cf.n_sections) {
                const double out = __dadd_rn(__dmul_rn(cf.c[s][0], v), z0[s]);
                z0[s] = __dadd_rn(__dadd_rn(__dmul_rn(cf.c[s][1], v), -__dmul_rn(cf.c[s][3], out)), z1[s]);
                z1[s] = __dadd_rn(__dmul_rn(cf.c[s][2], v), -__dmul_rn(cf.c[s][4], out));
                v = out;
            }
        }
        y[(long long)t * n_channels + c] = v;
    }
#pragma unroll
    for (int s = 0; s < kSosMaxSections; ++s) {
        if (s < cf.n_sections) { z[(2 * s) * n_channels + c] = z0[s]; z[(2 * s + 1) * n_channels + c] = z1[s]; }
    }
}

int sos_stream_run(const void* x, bool x_is_f64, int n, int n_channels, double* z, double* y, int first, int warm_start,
                   const SosCoefs& cf, cudaStream_t st) {
    ProfScope ps(kProfStream, st);
    const int grid = ceil_div(n_channels, 128);
    if (x_is_f64) k_sos_stream<double><<<grid, 128, 0, st>>>((const double*)x, n, n_channels, z, y, first, warm_start, cf);
    else k_sos_stream<float><<<grid, 128, 0, st>>>((const float*)x, n, n_channels, z, y, first, warm_start, cf);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// ---- Dequantization node: medians lookup + 5-tap smoothing across bins ---------------------------------------
__global__ void k_dequantize(const double* __restrict__ labels, const double* __restrict__ medians,
                             const double* __restrict__ taps, int radius, int smooth, int n_bins, int n_levels,
                             long long n_rows, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * n_bins) return;
    const long long r = idx / n_bins;
    const int b = (int)(idx - r * n_bins);
    auto raw = [&](int i) {
        if (i < 0) i = -i - 1;
        if (i >= n_bins) i = 2 * n_bins - 1 - i;
        int lv = (int)labels[r * n_bins + i];
        lv = lv < 0 ? 0 : (lv >= n_levels ? n_levels - 1 : lv);
        return medians[i * n_levels + lv];
    };
    double v = raw(b);
    if (smooth) {
        double t = __dmul_rn(v, taps[radius]);
        for (int jj = -radius; jj < 0; ++jj) t = __dadd_rn(t, __dmul_rn(__dadd_rn(raw(b + jj), raw(b - jj)), taps[radius + jj]));
        v = t;
    }
    out[idx] = v;
}

int dequantize_run(const double* labels, const double* medians, const double* taps, int radius, int smooth, int n_bins,
                   int n_levels, long long n_rows, double* out, cudaStream_t st) {
    const long long total = n_rows * n_bins;
    if (total == 0) return SGS_OK;
    k_dequantize<<<ceil_div(total, 128), 128, 0, st>>>(labels, medians, taps, radius, smooth, n_bins, n_levels, n_rows, out);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// ---- GriffinLim node: overlap-add of the newest block(s) with the ones still in the ring, low-pass, int16 ----
constexpr int kBlkLen = 480;

__global__ void __launch_bounds__(192)
k_gl_emit_stream(const double* __restrict__ block_ring /*[kBlockRing][480]*/, const double* __restrict__ ola_window,
                 double* __restrict__ lp_state, short* __restrict__ pcm, const __grid_constant__ LpCoefs c,
                 double norm_div, int first_frame, const __grid_constant__ EmitFrames fr) {
    __shared__ double v[kMaxFramesPerPush * 192];
    __shared__ int offs[kMaxFramesPerPush + 1];
    if (threadIdx.x == 0) {
        int o = 0;
        for (int q = 0; q < fr.n; ++q) { offs[q] = o; o += fr.pos[q] - fr.prev[q]; }
        offs[fr.n] = o;
    }
    __syncthreads();
    for (int q = 0; q < fr.n; ++q) {
        const int shifted = fr.pos[q] - fr.prev[q];
        const long long k = fr.index[q];
        for (int i = threadIdx.x; i < shifted; i += blockDim.x) {
            const int p = fr.pos[q] - kBlkLen + i;
            double num = 0.0, den = 0.0;
            for (long long jj = (k - 4 > first_frame ? k - 4 : first_frame); jj <= k; ++jj) {   // arrival order
                const int slot = (int)(jj & (kBlockRing - 1));
                if (fr.ring_index[slot] != jj) continue;
                const int off = p - (fr.ring_pos[slot] - kBlkLen);
                if (off >= 0 && off < kBlkLen) {
                    num += block_ring[slot * kBlkLen + off];
                    den += ola_window[off];
                }
            }
            v[offs[q] + i] = (den != 0.0) ? num / den : num;
        }
    }
    __syncthreads();
    const int total = offs[fr.n];
    if (threadIdx.x == 0) {
        // the recurrence itself is serial (2 dependent DFMA per sample); everything per-sample around it is not
        // run at the full width kLpMaxOrd with the coefficients beyond c.ord zero: the extra states stay 0 and the
        // live ones see the same operations (fma(b, x, 0) == b * x), so the state array stays in registers
        double z[kLpMaxOrd];
#pragma unroll
        for (int i = 0; i < kLpMaxOrd; ++i) z[i] = (i < c.ord) ? lp_state[i] : 0.0;
        for (int t = 0; t < total; ++t) {
            const double xin = v[t];
            const double y = fma(c.b[0], xin, z[0]);
#pragma unroll
            for (int i = 0; i < kLpMaxOrd - 1; ++i) z[i] = fma(-c.a[i + 1], y, fma(c.b[i + 1], xin, z[i + 1]));
            z[kLpMaxOrd - 1] = fma(-c.a[kLpMaxOrd], y, c.b[kLpMaxOrd] * xin);
            v[t] = y;
        }
#pragma unroll
        for (int i = 0; i < kLpMaxOrd; ++i) if (i < c.ord) lp_state[i] = z[i];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        double qv = v[t] / norm_div;
        qv = qv < -0.99 ? -0.99 : (qv > 0.99 ? 0.99 : qv);
        pcm[t] = (short)(int)(qv * 32767.0);
    }
}

int gl_emit_stream_run(const double* block_ring, const double* ola_window, double* lp_state, short* pcm, const LpCoefs& c,
                       double norm_div, int first_frame, const EmitFrames& fr, cudaStream_t st) {
    { ProfScope ps(kProfGlOla, st); k_gl_emit_stream<<<1, 192, 0, st>>>(block_ring, ola_window, lp_state, pcm, c, norm_div, first_frame, fr); }
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
