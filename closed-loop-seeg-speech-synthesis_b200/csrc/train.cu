// Training-side kernels (train.py:78-168 of the reference).
//
//   k_colminmax / k_quantize ... local/quantization.py:83-122 (logistic borders need per-bin min/max; labels =
//                                smallest interval whose border is >= the value)
//   ranking + correlation ....... scipy.stats.spearmanr per feature column against the frame-mean of the target
//                                (train.py:96-109): average ranks for ties, Pearson of the ranks
//   k_gram / k_class_sums ....... the data passes of sklearn LinearDiscriminantAnalysis(solver='svd').fit for all 40
//                                bins at once (train.py:112-118): every bin shares X, so the fit only needs
//                                G = Xc^T Xc, the per-(bin, class) sums of Xc and the class counts (closed form R5,
//                                SURVEY.md 8a'); the 150 x 150 eigen-problems are solved on the host.
// FP64 CUDA-core implementations; partial results are reduced in a fixed order (deterministic).
#include <math.h>
#include <vector>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <stdlib.h>
#include "kernels.cuh"

namespace sgs {

// ---- per-column min / max (one block per column) ---------------------------------------------------
__global__ void k_colminmax(const double* __restrict__ y, long long n, int ncol, double* __restrict__ mn, double* __restrict__ mx) {
    __shared__ double s_lo[32], s_hi[32];
    const int c = blockIdx.x;
    double lo = INFINITY, hi = -INFINITY;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { const double v = y[i * ncol + c]; lo = fmin(lo, v); hi = fmax(hi, v); }
    for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(~0u, lo, o)); hi = fmax(hi, __shfl_xor_sync(~0u, hi, o)); }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = threadIdx.x < (blockDim.x >> 5) ? s_lo[threadIdx.x] : INFINITY;
        hi = threadIdx.x < (blockDim.x >> 5) ? s_hi[threadIdx.x] : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(~0u, lo, o)); hi = fmax(hi, __shfl_xor_sync(~0u, hi, o)); }
        if (threadIdx.x == 0) { mn[c] = lo; mx[c] = hi; }
    }
}

__global__ void k_quantize(const double* __restrict__ y, const double* __restrict__ borders, long long n, int ncol, int nint,
                           double* __restrict__ q) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * ncol) return;
    const int c = (int)(idx % ncol);
    const double v = y[idx];
    int lab = 0;                                            // above every border: stays 0 (quantization.py:114)
    for (int k = nint - 1; k >= 0; --k)
        if (v <= borders[c * nint + k]) lab = k;
    q[idx] = (double)lab;
}

// ---- row mean of the target (np.mean(y_train, axis=1)) ------------------------------------------------
__global__ void k_rowmean(const double* __restrict__ y, long long n, int ncol, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int c = 0; c < ncol; ++c) s += y[i * ncol + c];
    out[i] = s / ncol;
}

// ---- transpose (n x ncol row-major) -> (ncol x n), adding an index payload for the sort --------------
__global__ void k_transpose(const double* __restrict__ x, long long n, int ncol, long long row_stride, double* __restrict__ xt) {
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + j; const int c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < n && c < ncol) ? x[r * row_stride + c] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j; const long long r = r0 + threadIdx.x;
        if (r < n && c < ncol) xt[(long long)c * n + r] = tile[threadIdx.x][j];
    }
}

__global__ void k_iota(int* __restrict__ idx, long long n, int nseg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * nseg) idx[i] = (int)(i % n);
}

// average ranks (1-based) of sorted keys, scattered back to the original positions.  -0.0 sorts before +0.0 in the
// radix order but compares equal, exactly the tie scipy.stats.rankdata sees.
__global__ void k_avg_ranks(const double* __restrict__ keys, const int* __restrict__ idx, long long n, int nseg,
                            double* __restrict__ ranks) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * nseg) return;
    const long long seg = g / n, i = g - seg * n;
    const double* k = keys + seg * n;
    const double v = k[i];
    long long lo = 0, hi = i;                               // first index with key == v
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (k[m] < v) lo = m + 1; else hi = m; }
    const long long first = lo;
    lo = i; hi = n;                                         // one past the last index with key == v
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (k[m] <= v) lo = m + 1; else hi = m; }
    ranks[seg * n + idx[g]] = 0.5 * (double)(first + lo - 1) + 1.0;
}

// Pearson correlation of rank vectors: one block per feature segment.  out[f] = rho, colsum[f] = sum of raw values.
__global__ void k_rank_corr(const double* __restrict__ rx /*[nseg][n]*/, const double* __restrict__ ry /*[n]*/,
                            const double* __restrict__ xt /*[nseg][n] raw*/, long long n, double* __restrict__ rho,
                            double* __restrict__ colsum) {
    __shared__ double red[4][32];
    const int f = blockIdx.x;
    const double mean = 0.5 * (double)(n + 1);
    double sxy = 0, sxx = 0, syy = 0, sraw = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double a = rx[(long long)f * n + i] - mean, b = ry[i] - mean;
        sxy = fma(a, b, sxy); sxx = fma(a, a, sxx); syy = fma(b, b, syy);
        sraw += xt[(long long)f * n + i];
    }
    double v[4] = {sxy, sxx, syy, sraw};
    for (int q = 0; q < 4; ++q) {
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(~0u, v[q], o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        for (int q = 0; q < 4; ++q) {
            double t = threadIdx.x < (blockDim.x >> 5) ? red[q][threadIdx.x] : 0.0;
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(~0u, t, o);
            v[q] = t;
        }
        if (threadIdx.x == 0) { rho[f] = v[0] / sqrt(v[1] * v[2]); colsum[f] = v[3]; }
    }
}

// ---- LDA sufficient statistics -------------------------------------------------------------------------
// column sums of the selected features (for the global mean), partial per slice then reduced
__global__ void k_colsum_partial(const double* __restrict__ x, const int* __restrict__ select, long long n, long long row_stride,
                                 int nf, int n_slices, double* __restrict__ part /*[slice][nf]*/) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (f >= nf) return;
    const long long r0 = n * s / n_slices, r1 = n * (s + 1) / n_slices;
    const int col = select[f];
    double acc = 0.0;
    for (long long r = r0; r < r1; ++r) acc += x[r * row_stride + col];
    part[(long long)s * nf + f] = acc;
}

__global__ void k_reduce_slices(const double* __restrict__ part, long long per_slice, int n_slices, double scale, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_slice) return;
    double acc = 0.0;
    for (int s = 0; s < n_slices; ++s) acc += part[(long long)s * per_slice + i];
    out[i] = acc * scale;
}

// G partial: 32x32 output tile per block over one row slice; Xc = X[:, select] - xbar formed on load
__global__ void __launch_bounds__(1024)
k_gram_partial(const double* __restrict__ x, const int* __restrict__ select, const double* __restrict__ xbar, long long n,
               long long row_stride, int nf, int n_slices, double* __restrict__ part /*[slice][nf][nf]*/) {
    __shared__ double A[32][33], B[32][33];
    const int ti = blockIdx.x, tj = blockIdx.y, s = blockIdx.z;
    if (tj < ti) return;                                    // symmetric: upper triangle only, mirrored at the reduce
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long r0 = n * s / n_slices, r1 = n * (s + 1) / n_slices;
    const int fa = ti * 32 + tx, fb = tj * 32 + tx;
    const int ca = fa < nf ? select[fa] : 0, cb = fb < nf ? select[fb] : 0;
    const double ma = fa < nf ? xbar[fa] : 0.0, mb = fb < nf ? xbar[fb] : 0.0;
    double acc = 0.0;
    for (long long r = r0; r < r1; r += 32) {
        const long long rr = r + ty;
        A[ty][tx] = (rr < r1 && fa < nf) ? x[rr * row_stride + ca] - ma : 0.0;
        B[ty][tx] = (rr < r1 && fb < nf) ? x[rr * row_stride + cb] - mb : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) acc = fma(A[k][ty], B[k][tx], acc);
        __syncthreads();
    }
    const int i = ti * 32 + ty, j = tj * 32 + tx;
    if (i < nf && j < nf) part[((long long)s * nf + i) * nf + j] = acc;
}

__global__ void k_gram_reduce(const double* __restrict__ part, int nf, int n_slices, double* __restrict__ G) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nf) return;
    const int a = i <= j ? i : j, b = i <= j ? j : i;       // tiles below the diagonal were skipped: read the mirror
    double acc = 0.0;
    for (int s = 0; s < n_slices; ++s) acc += part[((long long)s * nf + a) * nf + b];
    G[(long long)i * nf + j] = acc;
}

// class sums: block = (bin, slice); thread = feature; accumulators [n_classes][nf] in shared memory
__global__ void k_class_sums_partial(const double* __restrict__ x, const int* __restrict__ select, const double* __restrict__ xbar,
                                     const double* __restrict__ labels /*[n][n_bins]*/, long long n, long long row_stride, int nf,
                                     int n_bins, int n_classes, int n_slices, double* __restrict__ part /*[slice][bin][class][nf]*/,
                                     double* __restrict__ cnt_part /*[slice][bin][class]*/) {
    extern __shared__ double acc[];                         // [n_classes][nf] + counts [n_classes]
    const int b = blockIdx.x, s = blockIdx.y, f = threadIdx.x;
    double* cnt = acc + (size_t)n_classes * nf;
    for (int i = threadIdx.x; i < n_classes * nf + n_classes; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
    const long long r0 = n * s / n_slices, r1 = n * (s + 1) / n_slices;
    const int col = f < nf ? select[f] : 0;
    const double m = f < nf ? xbar[f] : 0.0;
    for (long long r = r0; r < r1; ++r) {
        int k = (int)labels[r * n_bins + b];
        k = k < 0 ? 0 : (k >= n_classes ? n_classes - 1 : k);
        if (f < nf) acc[k * nf + f] += x[r * row_stride + col] - m;
        if (f == 0) cnt[k] += 1.0;
    }
    __syncthreads();
    double* o = part + (((long long)s * n_bins + b) * n_classes) * nf;
    for (int i = threadIdx.x; i < n_classes * nf; i += blockDim.x) o[i] = acc[i];
    if (threadIdx.x < n_classes) cnt_part[((long long)s * n_bins + b) * n_classes + threadIdx.x] = cnt[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
int quantize_run(const double* y, long long n, int ncol, const double* borders, int nint, double* q, cudaStream_t st) {
    if (n == 0) return SGS_OK;
    k_quantize<<<ceil_div(n * ncol, 256), 256, 0, st>>>(y, borders, n, ncol, nint, q);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

int colminmax_run(const double* y, long long n, int ncol, double* mn, double* mx, cudaStream_t st) {
    k_colminmax<<<ncol, 256, 0, st>>>(y, n, ncol, mn, mx);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// Spearman rho of every column of x (n x ncol, row stride given) against the row mean of y (n x ny).
int spearman_run(const double* x, long long n, int ncol, long long row_stride, const double* y, int ny, double* rho,
                 double* colsum, cudaStream_t st) {
    // Columns are ranked in batches of ~64 M elements, each batch with the target as its last segment: the scratch (three
    // double and two int arrays per element + the sort's own) stays below 3 GB whatever the recording - all 640 columns of a
    // 1 h session at once took 10 GB, and growing the stream-ordered pool by that much on the first call cost 0.4 s, ten
    // times the ranking itself.  A column's result does not depend on the batch it is ranked in.
    SGS_ARG(n >= 1 && n < 1073741824LL, "bad row count %lld", n);
    const long long per_batch_elems = 64LL << 20;
    int batch = (int)(per_batch_elems / n);
    batch = batch < 1 ? 1 : (batch > ncol ? ncol : batch);
    SGS_ARG((long long)(batch + 1) * n < 2147483647LL, "too many elements for the segmented sort (%lld)", (long long)(batch + 1) * n);
    ProfScope ps(kProfTrain, st);
    const int nseg_max = batch + 1;                          // last segment of a batch = the target
    double *xt = nullptr, *keys = nullptr, *ranks = nullptr;
    int *idx_in = nullptr, *idx_out = nullptr, *offs = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    const size_t tot_max = (size_t)nseg_max * n;
    SGS_CUDA(cudaMallocAsync((void**)&xt, sizeof(double) * tot_max, st));
    SGS_CUDA(cudaMallocAsync((void**)&keys, sizeof(double) * tot_max, st));
    SGS_CUDA(cudaMallocAsync((void**)&ranks, sizeof(double) * tot_max, st));
    SGS_CUDA(cudaMallocAsync((void**)&idx_in, sizeof(int) * tot_max, st));
    SGS_CUDA(cudaMallocAsync((void**)&idx_out, sizeof(int) * tot_max, st));
    SGS_CUDA(cudaMallocAsync((void**)&offs, sizeof(int) * (nseg_max + 1), st));
    std::vector<int> h_offs(nseg_max + 1);
    for (int s = 0; s <= nseg_max; ++s) h_offs[s] = (int)((long long)s * n);
    SGS_CUDA(cudaMemcpyAsync(offs, h_offs.data(), sizeof(int) * (nseg_max + 1), cudaMemcpyHostToDevice, st));
    SGS_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp_bytes, xt, keys, idx_in, idx_out, (int)tot_max, nseg_max, offs, offs + 1, 0, 64, st));
    SGS_CUDA(cudaMallocAsync(&tmp, tmp_bytes, st));
    for (int c0 = 0; c0 < ncol; c0 += batch) {
        const int nb = ncol - c0 < batch ? ncol - c0 : batch, nseg = nb + 1;
        const size_t tot = (size_t)nseg * n;
        k_transpose<<<dim3(ceil_div(n, 32), ceil_div(nb, 32)), dim3(32, 8), 0, st>>>(x + c0, n, nb, row_stride, xt);
        SGS_LAUNCHED();
        k_rowmean<<<ceil_div(n, 256), 256, 0, st>>>(y, n, ny, xt + (size_t)nb * n);
        SGS_LAUNCHED();
        k_iota<<<ceil_div((long long)tot, 256), 256, 0, st>>>(idx_in, n, nseg);
        SGS_LAUNCHED();
        size_t need = tmp_bytes;
        SGS_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(tmp, need, xt, keys, idx_in, idx_out, (int)tot, nseg, offs, offs + 1, 0, 64, st));
        SGS_LAUNCHED();
        k_avg_ranks<<<ceil_div((long long)tot, 256), 256, 0, st>>>(keys, idx_out, n, nseg, ranks);
        SGS_LAUNCHED();
        k_rank_corr<<<nb, 256, 0, st>>>(ranks, ranks + (size_t)nb * n, xt, n, rho + c0, colsum + c0);
        SGS_LAUNCHED();
    }
    cudaFreeAsync(xt, st); cudaFreeAsync(keys, st); cudaFreeAsync(ranks, st); cudaFreeAsync(idx_in, st);
    cudaFreeAsync(idx_out, st); cudaFreeAsync(offs, st); cudaFreeAsync(tmp, st);
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// xbar[nf] (global mean of the selected columns), G[nf][nf] = Xc^T Xc, sums[bin][class][nf], counts[bin][class]
int col_means_run(const double* x, long long n, long long row_stride, const int* select, int nf, double* xbar, cudaStream_t st) {
    // enough row slices to fill the GPU (the decode path calls this on 1.9 M rows per step: 64 CTAs took 6.5 ms)
    int n_slices = (int)(n / 256);
    n_slices = n_slices < 1 ? 1 : (n_slices > 148 * 8 ? 148 * 8 : n_slices);
    double* p_col = nullptr;
    SGS_CUDA(cudaMallocAsync((void**)&p_col, sizeof(double) * n_slices * nf, st));
    k_colsum_partial<<<dim3(ceil_div(nf, 128), n_slices), 128, 0, st>>>(x, select, n, row_stride, nf, n_slices, p_col);
    SGS_LAUNCHED();
    k_reduce_slices<<<ceil_div(nf, 128), 128, 0, st>>>(p_col, nf, n_slices, 1.0 / (double)n, xbar);
    SGS_LAUNCHED();
    cudaFreeAsync(p_col, st);
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

int lda_stats_run(const double* x, long long n, long long row_stride, const int* select, int nf, const double* labels, int n_bins,
                  int n_classes, double* xbar, double* G, double* sums, double* counts, const double* xbar_in, cudaStream_t st) {
    ProfScope ps(kProfTrain, st);
    int n_slices = (int)(n / 2048);
    n_slices = n_slices < 1 ? 1 : (n_slices > 64 ? 64 : n_slices);
    const int nt = ceil_div(nf, 32);
    double *p_col = nullptr, *p_g = nullptr, *p_s = nullptr, *p_c = nullptr;
    SGS_CUDA(cudaMallocAsync((void**)&p_col, sizeof(double) * n_slices * nf, st));
    SGS_CUDA(cudaMallocAsync((void**)&p_g, sizeof(double) * (size_t)n_slices * nf * nf, st));
    SGS_CUDA(cudaMallocAsync((void**)&p_s, sizeof(double) * (size_t)n_slices * n_bins * n_classes * nf, st));
    SGS_CUDA(cudaMallocAsync((void**)&p_c, sizeof(double) * (size_t)n_slices * n_bins * n_classes, st));
    if (xbar_in) {
        SGS_CUDA(cudaMemcpyAsync(xbar, xbar_in, sizeof(double) * nf, cudaMemcpyDeviceToDevice, st));
    } else {
        k_colsum_partial<<<dim3(ceil_div(nf, 128), n_slices), 128, 0, st>>>(x, select, n, row_stride, nf, n_slices, p_col);
        SGS_LAUNCHED();
        k_reduce_slices<<<ceil_div(nf, 128), 128, 0, st>>>(p_col, nf, n_slices, 1.0 / (double)n, xbar);
        SGS_LAUNCHED();
    }
    const char* env_tc = getenv("SGS_TRAIN_TC");
    if (lda_stats_tc_supported(nf, n_bins, n_classes) && !(env_tc && env_tc[0] == '0')) {
        // tensor-core path (train_tc.cu): exact integer GEMMs on int8 digit matrices
        cudaFreeAsync(p_col, st); cudaFreeAsync(p_g, st); cudaFreeAsync(p_s, st); cudaFreeAsync(p_c, st);
        return lda_stats_tc_run(x, n, row_stride, select, nf, labels, n_bins, n_classes, xbar, G, sums, counts, st);
    }
    SGS_CUDA(cudaMemsetAsync(p_g, 0, sizeof(double) * (size_t)n_slices * nf * nf, st));
    k_gram_partial<<<dim3(nt, nt, n_slices), dim3(32, 32), 0, st>>>(x, select, xbar, n, row_stride, nf, n_slices, p_g);
    SGS_LAUNCHED();
    k_gram_reduce<<<dim3(ceil_div(nf, 128), nf), 128, 0, st>>>(p_g, nf, n_slices, G);
    SGS_LAUNCHED();
    const size_t smem = sizeof(double) * ((size_t)n_classes * nf + n_classes);
    const int threads = ((nf + 31) / 32) * 32;
    k_class_sums_partial<<<dim3(n_bins, n_slices), threads, smem, st>>>(x, select, xbar, labels, n, row_stride, nf, n_bins,
                                                                       n_classes, n_slices, p_s, p_c);
    SGS_LAUNCHED();
    const long long per = (long long)n_bins * n_classes * nf;
    k_reduce_slices<<<ceil_div(per, 128), 128, 0, st>>>(p_s, per, n_slices, 1.0, sums);
    SGS_LAUNCHED();
    k_reduce_slices<<<ceil_div((long long)n_bins * n_classes, 128), 128, 0, st>>>(p_c, (long long)n_bins * n_classes, n_slices, 1.0, counts);
    SGS_LAUNCHED();
    cudaFreeAsync(p_col, st); cudaFreeAsync(p_g, st); cudaFreeAsync(p_s, st); cudaFreeAsync(p_c, st);
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
