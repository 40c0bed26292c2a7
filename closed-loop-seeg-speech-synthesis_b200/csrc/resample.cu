// Audio decimation on the device: scipy.signal.decimate(audio, q) as train.py:125 calls it (48 kHz -> 16 kHz before the
// log-mel target) = zero-phase IIR low-pass + keep every q-th sample.
//
// Reference semantics restated (scipy.signal.decimate, ftype='iir', zero_phase=True, as shipped in the oracle's scipy):
//   sos = cheby1(8, 0.05, 0.8 / q, output='sos')                          (host, sgs/spectrogram.py)
//   y   = sosfiltfilt(sos, x)[::q]:  ext = odd extension of x by edge = 3 * (2 * n_sections + 1) samples at both ends;
//         forward sosfilt from zi * ext[0]; backward sosfilt (on the reversed result) from zi * (its first sample);
//         reverse, drop the extension.
// Both passes are linear recurrences over ONE long sequence (57.6 M .. 173 M samples for an hour of audio), so time is the
// only axis to parallelise: the sequence is cut into chunks, one thread per chunk; a chunk other than the first starts
// `warm` samples early from the zero state, by which time the influence of the true state has decayed below 2^-70 (the
// host derives `warm` from the largest pole radius; 910 samples for q = 3).  The first chunk of each pass starts from the
// exact scipy initial state.  The backward pass writes only the samples decimation keeps.
#include <math.h>
#include "kernels.cuh"

namespace sgs {

constexpr int kDecMaxSections = 8;
struct DecimCoefs {
    double c[kDecMaxSections][5];        // b0 b1 b2 a1 a2
    double zi[kDecMaxSections][2];       // sosfilt_zi
    int n_sections;
};

__device__ __forceinline__ double odd_ext(const double* __restrict__ x, long long n, int edge, long long i) {
    if (i < edge) return 2.0 * x[0] - x[edge - i];
    if (i < edge + n) return x[i - edge];
    return 2.0 * x[n - 1] - x[n - 2 - (i - (edge + n))];
}

// PASS 0: forward over ext -> tmp[N].  PASS 1: backward over tmp -> out[j] for every kept sample.
template <int PASS>
__global__ void __launch_bounds__(128)
k_sosfiltfilt(const double* __restrict__ x, long long n, int edge, double* __restrict__ tmp, double* __restrict__ out, int q,
              long long chunk, int warm, const __grid_constant__ DecimCoefs cf) {
    const long long N = n + 2LL * edge;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r_lo = c * chunk;
    if (r_lo >= N) return;
    const long long r_hi = r_lo + chunk < N ? r_lo + chunk : N;
    const long long r0 = (c == 0) ? 0 : (r_lo - warm > 0 ? r_lo - warm : 0);
    auto sample = [&](long long r) -> double { return PASS == 0 ? odd_ext(x, n, edge, r) : tmp[N - 1 - r]; };
    double z0[kDecMaxSections], z1[kDecMaxSections];
    const double first = (r0 == 0) ? sample(0) : 0.0;
#pragma unroll
    for (int s = 0; s < kDecMaxSections; ++s) {
        z0[s] = (r0 == 0 && s < cf.n_sections) ? cf.zi[s][0] * first : 0.0;     // scipy scales every section's zi by the first INPUT sample
        z1[s] = (r0 == 0 && s < cf.n_sections) ? cf.zi[s][1] * first : 0.0;
    }
    for (long long r = r0; r < r_hi; ++r) {
        double v = sample(r);
#pragma unroll
        for (int s = 0; s < kDecMaxSections; ++s) {
            if (s < cf.n_sections) {
                // scipy's _sosfilt: y = b0 x + z0; z0 = b1 x - a1 y + z1; z1 = b2 x - a2 y (separate multiplies and adds)
                const double y = __dadd_rn(__dmul_rn(cf.c[s][0], v), z0[s]);
                z0[s] = __dadd_rn(__dadd_rn(__dmul_rn(cf.c[s][1], v), -__dmul_rn(cf.c[s][3], y)), z1[s]);
                z1[s] = __dadd_rn(__dmul_rn(cf.c[s][2], v), -__dmul_rn(cf.c[s][4], y));
                v = y;
            }
        }
        if (r >= r_lo) {
            if (PASS == 0) tmp[r] = v;
            else {
                const long long t = (N - 1 - r) - edge;                          // position in the un-extended signal
                if (t >= 0 && t < n && t % q == 0) out[t / q] = v;
            }
        }
    }
}

int decimate_run(const double* x, long long n, int q, const DecimCoefs& cf, int edge, int warm, double* out, cudaStream_t st) {
    const long long N = n + 2LL * edge;
    double* tmp = nullptr;
    SGS_CUDA(cudaMallocAsync((void**)&tmp, sizeof(double) * N, st));
    long long chunk = 8LL * warm;                                               // 12.5 % redundant work
    if (chunk < 1024) chunk = 1024;
    const long long n_chunks = (N + chunk - 1) / chunk;
    const int grid = ceil_div(n_chunks, 128);
    { ProfScope ps(kProfLogMel, st); k_sosfiltfilt<0><<<grid, 128, 0, st>>>(x, n, edge, tmp, out, q, chunk, warm, cf); }
    SGS_LAUNCHED();
    { ProfScope ps(kProfLogMel, st); k_sosfiltfilt<1><<<grid, 128, 0, st>>>(x, n, edge, tmp, out, q, chunk, warm, cf); }
    SGS_LAUNCHED();
    cudaFreeAsync(tmp, st);
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs

#include "../../include/sgs.h"

extern "C" int sgs_decimate(const double* audio, int64_t n, int q, const double* sos, const double* zi, int n_sections, int edge,
                            int warm, double* out, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(audio && sos && zi && out, "NULL argument");
    SGS_ARG(q >= 1 && n_sections >= 1 && n_sections <= kDecMaxSections && edge >= 0 && warm >= 1, "bad arguments");
    SGS_ARG(n > edge, "the signal (%lld samples) must be longer than the padding (%d)", (long long)n, edge);
    DecimCoefs cf;
    memset(&cf, 0, sizeof(cf));
    cf.n_sections = n_sections;
    for (int s = 0; s < n_sections; ++s) {
        SGS_ARG(sos[s * 6 + 3] == 1.0, "sos section %d is not normalised (a0 = %g)", s, sos[s * 6 + 3]);
        cf.c[s][0] = sos[s * 6 + 0]; cf.c[s][1] = sos[s * 6 + 1]; cf.c[s][2] = sos[s * 6 + 2];
        cf.c[s][3] = sos[s * 6 + 4]; cf.c[s][4] = sos[s * 6 + 5];
        cf.zi[s][0] = zi[s * 2]; cf.zi[s][1] = zi[s * 2 + 1];
    }
    const int64_t n_out = (n + q - 1) / q;
    Staged sa, so;
    int rc = stage_in(sa, audio, sizeof(double) * (size_t)n, st);
    if (rc == SGS_OK) rc = stage_out(so, out, sizeof(double) * (size_t)n_out, st);
    if (rc == SGS_OK) rc = decimate_run((const double*)sa.dev, n, q, cf, edge, warm, (double*)so.dev, st);
    if (rc == SGS_OK) rc = finish_out(so, st);
    const bool sync = so.host != nullptr;
    release(sa, st); release(so, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}
