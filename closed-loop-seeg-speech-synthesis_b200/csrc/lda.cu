// Per-mel-bin LDA decoding fused with dequantisation and the 5-tap smoothing.
//
// Reference semantics restated (paths under the reference tree):
//   livenodes/LDASynthesis.py:19-28   per bin: scores = x[select] . coef_^T + intercept_; label = classes_[argmax]
//                                     (sklearn: first maximum wins; binary estimators: single score > 0)
//   livenodes/Dequantization.py:15-18 spec = medians[bin, label] then scipy.ndimage.gaussian_filter(sigma=0.5)
//                                     across the bins, mode 'reflect' (the batch helper quantization.py:125-135
//                                     does not smooth)
//
// FP64 CUDA-core kernel.  The stacked feature vector of a frame is never materialised: feature f of row r is
// gathered straight from the un-stacked log-power array, feat[r + first_row - (order - tap)*step][c] with
// (c, tap) = divmod(select[f], order + 1), exact zero before the stream start.
//
// Layout: W is repacked on upload to [bin][feature][class] so that the 9 class weights of one (bin, feature)
// are contiguous and warp-uniform.  Block = 4 warps x 32 frames (lane = frame); warp w scores bins w, w+4, ...
#include <math.h>
#include <algorithm>
#include "kernels.cuh"

namespace sgs {

constexpr int kLdaFrames = 32;
constexpr int kLdaWarps = 4;
constexpr int kMaxClasses = 9;


template <int KC>
__global__ void __launch_bounds__(kLdaFrames * kLdaWarps)
k_lda_decode(const double* __restrict__ feat, const double* __restrict__ Wt /*[bin][F][KC]*/,
             const double* __restrict__ bias /*[bin][KC]*/, const double* __restrict__ cls /*[bin][KC]*/,
             const int* __restrict__ select, const double* __restrict__ medians, const double* __restrict__ taps,
             double* __restrict__ labels, double* __restrict__ spec, int smooth, const LdaGeom g,
             const int* __restrict__ list, const int* __restrict__ list_count) {
    extern __shared__ double sm[];
    double* xs = sm;                                   // [F][33]
    double* raw = sm + (size_t)g.n_features * 33;      // [n_bins][33]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // list mode (exact re-scoring of the frames the tensor-core pass flagged): a fixed grid strides over the list, whose
    // length is only known on the device; otherwise one pass with vb = blockIdx.x
    const int n_list = list ? *list_count : 0;
    for (int vb = blockIdx.x; list ? vb * kLdaFrames < n_list : vb == (int)blockIdx.x; vb += gridDim.x) {
    int sess = blockIdx.y;
    int row = vb * kLdaFrames + lane;
    bool live = row < g.n_rows;
    if (list) {
        const int i = vb * kLdaFrames + lane;                   // frame id = session * n_rows + row
        live = i < n_list;
        const int id = live ? list[i] : list[n_list - 1];
        sess = id / g.n_rows;
        row = id - sess * g.n_rows;
    }
    const double* fs = feat + (long long)sess * g.n_windows * g.n_channels;

    // gather the selected, stacked features of 32 frames
    for (int f = warp; f < g.n_features; f += kLdaWarps) {
        const int col = select[f];
        const int c = col / (g.order + 1), tap = col - c * (g.order + 1);
        const int w = row + g.first_row - (g.order - tap) * g.step;
        xs[f * 33 + lane] = (live && w >= 0) ? fs[(long long)w * g.n_channels + c] : 0.0;
    }
    __syncthreads();

    for (int b = warp; b < g.n_bins; b += kLdaWarps) {
        double acc[KC];
        const double* bb = bias + b * KC;
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[k] = 0.0;
        const double* wb = Wt + (long long)b * g.n_features * KC;
        for (int f = 0; f < g.n_features; ++f) {
            const double xv = xs[f * 33 + lane];
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[k] = fma(xv, __ldg(wb + f * KC + k), acc[k]);
        }
        int best = 0;
        double bv = acc[0] + bb[0];
#pragma unroll
        for (int k = 1; k < KC; ++k) {
            const double s = acc[k] + bb[k];
            if (s > bv) { bv = s; best = k; }              // strict: the first maximum wins, as numpy argmax
        }
        const double lab = cls[b * KC + best];
        if (live && labels) labels[((long long)sess * g.n_rows + row) * g.n_bins + b] = lab;
        int lv = (int)lab;
        lv = lv < 0 ? 0 : (lv >= g.n_levels ? g.n_levels - 1 : lv);
        raw[b * 33 + lane] = medians[b * g.n_levels + lv];
    }
    if (spec) {
    __syncthreads();
    for (int b = warp; b < g.n_bins; b += kLdaWarps) {
        double v = raw[b * 33 + lane];
        if (smooth) {
            // scipy.ndimage correlate1d, symmetric-kernel branch: centre tap first, then pairs from the far end in;
            // separate multiply and add (no FMA) so the result is bit-identical to the C loop
            const int R = g.smooth_radius;
            auto at = [&](int i) {                          // 'reflect': d c b a | a b c d | d c b a
                if (i < 0) i = -i - 1;
                if (i >= g.n_bins) i = 2 * g.n_bins - 1 - i;
                return raw[i * 33 + lane];
            };
            double t = __dmul_rn(v, taps[R]);
            for (int jj = -R; jj < 0; ++jj) t = __dadd_rn(t, __dmul_rn(__dadd_rn(at(b + jj), at(b - jj)), taps[R + jj]));
            v = t;
        }
        if (live) spec[((long long)sess * g.n_rows + row) * g.n_bins + b] = v;
    }
    }
    __syncthreads();                                            // xs / raw are reused by the next list block
    }
}

// ---- few frames (the streaming nodes): one block per frame, one thread per (bin, class) ---------------------------
// Same arithmetic per score as k_lda_decode - acc = fma(x_f, W[b][f][k], acc) for f ascending, then + bias, strict
// argmax in class order - so a frame decodes to the same bits whichever kernel runs; only the mapping differs: the
// 360 scores of a frame are 360 independent 150-long FMA chains instead of 40 x 9 chains walked by one thread.
template <int KC>
__global__ void __launch_bounds__(512)
k_lda_rows(const double* __restrict__ feat, const double* __restrict__ Wt, const double* __restrict__ bias,
           const double* __restrict__ cls, const int* __restrict__ select, const double* __restrict__ medians,
           const double* __restrict__ taps, double* __restrict__ labels, double* __restrict__ spec, int smooth, const LdaGeom g,
           const int* __restrict__ list, const int* __restrict__ list_count) {
    extern __shared__ double sm[];
    double* xs = sm;                                   // [F]
    double* score = xs + g.n_features;                 // [n_bins][KC]
    double* raw = score + g.n_bins * KC;               // [n_bins]
    // list mode (exact re-scoring of the frames the tensor-core pass flagged; frame id = session * n_rows + row): a
    // fixed grid strides over the list, whose length is only known on the device
    const int n_list = list ? *list_count : 1;
    for (int li = list ? blockIdx.x : 0; li < n_list; li += list ? gridDim.x : n_list) {
    int row = blockIdx.x, sess = blockIdx.y;
    if (list) { const int id = list[li]; sess = id / g.n_rows; row = id - sess * g.n_rows; }
    const double* fs = feat + (long long)sess * g.n_windows * g.n_channels;
    for (int f = threadIdx.x; f < g.n_features; f += blockDim.x) {
        const int col = select[f];
        const int c = col / (g.order + 1), tap = col - c * (g.order + 1);
        const int w = row + g.first_row - (g.order - tap) * g.step;
        xs[f] = (w >= 0) ? fs[(long long)w * g.n_channels + c] : 0.0;
    }
    __syncthreads();
    const int n_scores = g.n_bins * KC;
    for (int i = threadIdx.x; i < n_scores; i += blockDim.x) {
        const int b = i / KC, k = i - b * KC;
        const double* wb = Wt + (long long)b * g.n_features * KC + k;
        double acc = 0.0;
#pragma unroll 10
        for (int f = 0; f < g.n_features; ++f) acc = fma(xs[f], __ldg(wb + f * KC), acc);
        score[i] = acc + bias[i];
    }
    __syncthreads();
    for (int b = threadIdx.x; b < g.n_bins; b += blockDim.x) {
        int best = 0;
        double bv = score[b * KC];
#pragma unroll
        for (int k = 1; k < KC; ++k) {
            const double s = score[b * KC + k];
            if (s > bv) { bv = s; best = k; }
        }
        const double lab = cls[b * KC + best];
        if (labels) labels[((long long)sess * g.n_rows + row) * g.n_bins + b] = lab;
        int lv = (int)lab;
        lv = lv < 0 ? 0 : (lv >= g.n_levels ? g.n_levels - 1 : lv);
        raw[b] = medians[b * g.n_levels + lv];
    }
    __syncthreads();
    for (int b = threadIdx.x; spec && b < g.n_bins; b += blockDim.x) {
        double v = raw[b];
        if (smooth) {
            const int R = g.smooth_radius;
            auto at = [&](int i) {
                if (i < 0) i = -i - 1;
                if (i >= g.n_bins) i = 2 * g.n_bins - 1 - i;
                return raw[i];
            };
            double t = __dmul_rn(v, taps[R]);
            for (int jj = -R; jj < 0; ++jj) t = __dadd_rn(t, __dmul_rn(__dadd_rn(at(b + jj), at(b - jj)), taps[R + jj]));
            v = t;
        }
        spec[((long long)sess * g.n_rows + row) * g.n_bins + b] = v;
    }
    __syncthreads();                                    // xs / score / raw are reused by the next list entry
    }
}

// ---- exact re-scoring of the (frame, bin) pairs the tensor-core filter could not decide (lda_tc.cu) -----------------------------
// list[e] = frame * 3 + slice of a flagged entry, flags[list[e]] = bit mask of its undecided bins (relative to the slice's
// first bin).  One warp per entry: the frame's selected, stacked features are gathered once into shared memory; lanes
// (p, k) = (lane / 9, lane % 9), p < 3, score class k of the p-th undecided bin with the arithmetic of k_lda_decode -
// acc = fma(x_f, W[b][f][k], acc) for f ascending, then + bias, strict argmax in class order - so the label equals the
// fp64 kernel's to the bit.
constexpr int kPairWarps = 8;
template <int KC>
__global__ void __launch_bounds__(kPairWarps * 32)
k_lda_pairs(const double* __restrict__ feat, const double* __restrict__ Wt, const double* __restrict__ bias,
            const double* __restrict__ cls, const int* __restrict__ select, double* __restrict__ labels, const LdaGeom g,
            const int* __restrict__ flags, const int* __restrict__ list, const int* __restrict__ list_count,
            const int* __restrict__ slice_bins) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xs = sm + (size_t)warp * g.n_features;
    const int n_list = *list_count;
    const int p = lane / KC, k = lane - p * KC;
    for (int e = blockIdx.x * kPairWarps + warp; e < n_list; e += gridDim.x * kPairWarps) {
        const int entry = list[e];
        const int frame = entry / 3, slice = entry - 3 * frame;
        unsigned mask = (unsigned)flags[entry];
        const int bin0 = slice_bins[slice];
        const int sess = frame / g.n_rows, row = frame - sess * g.n_rows;
        const double* fs = feat + (long long)sess * g.n_windows * g.n_channels;
        for (int f = lane; f < g.n_features; f += 32) {
            const int col = select[f];
            const int c = col / (g.order + 1), tap = col - c * (g.order + 1);
            const int w = row + g.first_row - (g.order - tap) * g.step;
            xs[f] = (w >= 0) ? fs[(long long)w * g.n_channels + c] : 0.0;
        }
        __syncwarp();
        while (mask) {
            int myb = -1;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (mask) {
                    const int bb = __ffs(mask) - 1;
                    if (q == p) myb = bb;
                    mask &= mask - 1;
                }
            }
            const bool active = myb >= 0;                                // p == 3 (lanes 27-31) never matches
            const int b = bin0 + (active ? myb : 0);
            double s = -INFINITY;
            if (active) {
                const double* wb = Wt + (long long)b * g.n_features * KC + k;
                double acc = 0.0;
#pragma unroll 10
                for (int f = 0; f < g.n_features; ++f) acc = fma(xs[f], __ldg(wb + f * KC), acc);
                s = acc + bias[b * KC + k];
            }
            const int base = p < 3 ? p * KC : 0;
            int best = 0;
            double bv = __shfl_sync(0xffffffffu, s, base);
#pragma unroll
            for (int j = 1; j < KC; ++j) {
                const double sj = __shfl_sync(0xffffffffu, s, base + j);
                if (sj > bv) { bv = sj; best = j; }                      // strict: the first maximum wins, as numpy argmax
            }
            if (active && k == 0) labels[(long long)frame * g.n_bins + b] = cls[b * KC + best];
        }
        __syncwarp();                                                    // xs is reused by the next entry
    }
}

int lda_pairs_run(const double* feat, const double* Wt, const double* bias, const double* cls, const int* select, double* labels,
                  const LdaGeom& g, cudaStream_t st, const int* flags, const int* list, const int* list_count, const int* slice_bins,
                  long long list_cap) {
    if (g.n_classes != kMaxClasses) { set_error("LDA kernel is built for %d classes per bin (got %d)", kMaxClasses, g.n_classes); return SGS_ERR_UNSUPPORTED; }
    const size_t smem = sizeof(double) * kPairWarps * (size_t)g.n_features;
    if (smem > 48 * 1024) { set_error("too many features (%d) for the pair re-scoring kernel", g.n_features); return SGS_ERR_UNSUPPORTED; }
    const long long want = (list_cap + kPairWarps - 1) / kPairWarps;
    const int grid = (int)std::min<long long>(std::max<long long>(want, 1), 148 * 8);
    { ProfScope ps(kProfLda, st); k_lda_pairs<kMaxClasses><<<grid, kPairWarps * 32, smem, st>>>(feat, Wt, bias, cls, select, labels, g, flags, list, list_count, slice_bins); }
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

constexpr long long kLdaRowsMax = 256;      // up to this many frames the per-frame kernel has the lower latency

int lda_run(const double* feat, const double* Wt, const double* bias, const double* cls, const int* select,
            const double* medians, const double* taps, double* labels, double* spec, int smooth, int n_sessions,
            const LdaGeom& g, cudaStream_t st, const int* list, const int* list_count, long long list_cap) {
    if (g.n_rows <= 0) return SGS_OK;
    const size_t smem = sizeof(double) * 33 * ((size_t)g.n_features + g.n_bins);
    if (smem > 200 * 1024) { set_error("too many features (%d) for the LDA kernel", g.n_features); return SGS_ERR_UNSUPPORTED; }
    if (g.n_classes != kMaxClasses) { set_error("LDA kernel is built for %d classes per bin (got %d)", kMaxClasses, g.n_classes); return SGS_ERR_UNSUPPORTED; }
    if (list || (long long)g.n_rows * n_sessions <= kLdaRowsMax) {
        // few frames, or the flagged near-ties of the tensor-core pass (a few dozen of 1.9 M frames in the bench: with 32
        // frames per block only two blocks of the batch kernel had work, 0.6 ms of exposed L2 latency)
        const int threads = 384;
        const size_t sm_rows = sizeof(double) * ((size_t)g.n_features + (size_t)g.n_bins * (kMaxClasses + 1));
        if (sm_rows > 48 * 1024) { set_error("too many features (%d) for the per-frame LDA kernel", g.n_features); return SGS_ERR_UNSUPPORTED; }
        const dim3 grid_rows = list ? dim3((unsigned)std::min<long long>(std::max<long long>(list_cap, 1), 148 * 4), 1) : dim3(g.n_rows, n_sessions);
        { ProfScope ps(kProfLda, st); k_lda_rows<kMaxClasses><<<grid_rows, threads, sm_rows, st>>>(feat, Wt, bias, cls, select, medians, taps, labels, spec, smooth, g, list, list_count); }
        SGS_LAUNCHED();
        SGS_CUDA(cudaGetLastError());
        return SGS_OK;
    }
    static unsigned long long optin = 0;
    SGS_CUDA(smem_optin(k_lda_decode<kMaxClasses>, 200 * 1024, &optin));
    dim3 grid(ceil_div(g.n_rows, kLdaFrames), n_sessions);
    { ProfScope ps(kProfLda, st); k_lda_decode<kMaxClasses><<<grid, kLdaFrames * kLdaWarps, smem, st>>>(feat, Wt, bias, cls, select, medians, taps, labels, spec, smooth, g, list, list_count); }
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
