// LDA training statistics on the 5th-generation tensor cores: the Gram matrix G = Xc^T Xc and the per-(bin, class)
// sums of Xc are dense contractions over the rows (time), so they run as tcgen05 GEMMs - in EXACT integer arithmetic.
//
// Reference: train.py:112-118 fits 40 sklearn LinearDiscriminantAnalysis(solver='svd') models on one shared X; closed
// form R5 (SURVEY.md 8a') needs only G, the class sums and the class counts.  sklearn works in float64 and tcgen05 has
// no fp64 kind, so the operands are split into int8 digits and multiplied with kind::i8 (s8 x s8 -> s32 accumulators in
// TMEM), an Ozaki-style error-free scheme:
//
//   v[r][f] = x[r][select f] - xbar[f]                      fp64, as the fp64 kernels form it
//   q[r][f] = rint(v * 2^(46 - e_f)),  2^e_f > max_r |v|    46-bit fixed point per feature (the only rounding made)
//   q = sum_i d_i 256^i,  d_i in [-128, 127],  i < 6        balanced base-256 digits: six int8 matrices D_0 .. D_5
//   Xq^T Xq = sum_{i,j} 256^(i+j) D_i^T D_j                 every D_i^T D_j is an integer GEMM; all 21 pairs i <= j are
//                                                           computed, so G is the exact Gram matrix of the quantised data
//   class sums = sum_j 256^j OneHot^T D_j                   OneHot[r][bin*9 + class] in {0, 1} is exact in int8 too
//
// int32 accumulators cannot overflow: a CTA contracts at most 65 536 rows (128 * 128 * 2^16 = 2^30); partial results are
// added into 64-bit integer accumulators in HBM with atomics (integer addition: order-independent, bit-reproducible),
// and only the last step - sum_c 2^(8c) T_c in ascending c and the power-of-two feature scales - is floating point.
//
// k_tc_absmax   per-feature max |v| (sets e_f)
// k_tc_pack     digits and one-hot labels written to HBM directly in the canonical K-major core-matrix layout the MMA
//               reads: [digit][16-row group][160 columns][16 B] and [16-row group][384 columns][16 B]; class counts
// k_tc_stats    grid (row splits, 27 roles): 21 Gram roles (digit pair i <= j: A = D_i columns [0,128) and [32,160) as two
//               M = 128 tiles, B = D_j, N = 160) and 6 class-sum roles (A = three 128-column tiles of OneHot, B = D_j).
//               warp 0 lane 0: TMA producer (cp.async.bulk of 128-row chunks into a 3-stage ring, mbarrier complete_tx);
//               warp 1 lane 0: MMA issuer (4 tcgen05.mma K = 32 per chunk and accumulator, tcgen05.commit frees the stage);
//               all 4 warps: epilogue, tcgen05.ld -> 64-bit atomic adds.
// k_tc_combine  integer accumulators -> G, class sums, counts in fp64.
#include <math.h>
#include "kernels.cuh"
#include "tc.cuh"

namespace sgs {

constexpr int kTrF = 160;                   // feature columns per digit matrix (model features padded with zeros)
constexpr int kTrOH = 384;                  // one-hot columns (bins * classes padded)
constexpr int kTrDigits = 6;
constexpr int kTrFixed = 46;                // fixed-point bits below 2^e_f
constexpr int kTrChunkRows = 128, kTrGroups = kTrChunkRows / 16;
constexpr uint32_t kTrLboD = kTrF * 16, kTrLboO = kTrOH * 16, kTrSbo = 128;
constexpr int kTrDChunk = kTrGroups * kTrLboD;          // 20 480 B: one digit matrix, 128 rows
constexpr int kTrOChunk = kTrGroups * kTrLboO;          // 49 152 B: one-hot, 128 rows
constexpr int kTrStage = kTrOChunk + kTrDChunk;         // 69 632 B (class-sum roles; Gram roles use 2 x 20 480)
constexpr int kTrStages = 3;
constexpr int kTrPairs = kTrDigits * (kTrDigits + 1) / 2;               // 21
constexpr int kTrRoles = kTrPairs + kTrDigits;                          // 27
constexpr int kTrAccPerRole = 3;
constexpr int kTrSmem = kTrStages * kTrStage + 256;
constexpr int kTrMaxChunksPerSplit = 65536 / kTrChunkRows;              // int32 headroom (see above)
constexpr long long kTrAccElems = 128LL * kTrF;                         // one accumulator: 128 rows x 160 columns

__host__ __device__ inline void tr_pair(int role, int& i, int& j) {     // role -> digit pair, i <= j, row-major enumeration
    int r = role;
    for (i = 0; i < kTrDigits; ++i) {
        const int n = kTrDigits - i;
        if (r < n) { j = i + r; return; }
        r -= n;
    }
    i = j = 0;
}

// ---- max |x - xbar| per model feature (non-negative doubles order like their bit patterns) -------------------------
__global__ void k_tc_absmax(const double* __restrict__ x, const int* __restrict__ select, const double* __restrict__ xbar, long long n,
                            long long row_stride, int nf, unsigned long long* __restrict__ amax_bits) {
    const int f = threadIdx.x;
    if (f >= nf) return;
    const int col = select[f];
    const double m = xbar[f];
    const long long r0 = n * blockIdx.x / gridDim.x, r1 = n * (blockIdx.x + 1) / gridDim.x;
    double a = 0.0;
    for (long long r = r0; r < r1; ++r) a = fmax(a, fabs(x[r * row_stride + col] - m));
    atomicMax(amax_bits + f, (unsigned long long)__double_as_longlong(a));
}

__global__ void k_tc_scales(const unsigned long long* __restrict__ amax_bits, int nf, double* __restrict__ to_fixed /*2^(46-e)*/,
                            double* __restrict__ from_fixed /*2^(e-46)*/) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= kTrF) return;
    int e = 0;
    if (f < nf) frexp(__longlong_as_double((long long)amax_bits[f]), &e);        // amax < 2^e (e = 0 for a constant column)
    to_fixed[f] = ldexp(1.0, kTrFixed - e);
    from_fixed[f] = ldexp(1.0, e - kTrFixed);
}

// ---- pack: block = one 128-row chunk; threads 0..159 = feature columns, 160..160+n_bins-1 = label bins ---------------
__global__ void __launch_bounds__(256)
k_tc_pack(const double* __restrict__ x, const int* __restrict__ select, const double* __restrict__ xbar,
          const double* __restrict__ to_fixed, const double* __restrict__ labels, long long n, long long row_stride, int nf,
          int n_bins, int n_classes, long long n_groups /*padded*/, unsigned char* __restrict__ digits, unsigned char* __restrict__ onehot,
          unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int s_cnt[];                 // [n_bins * n_classes]
    const int t = threadIdx.x;
    for (int i = t; i < n_bins * n_classes; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const long long g0 = (long long)blockIdx.x * kTrGroups;
    if (t < kTrF) {
        const bool live = t < nf;
        const int col = live ? select[t] : 0;
        const double m = live ? xbar[t] : 0.0, sc = live ? to_fixed[t] : 0.0;
        for (int g = 0; g < kTrGroups; ++g) {
            const long long r0 = (g0 + g) * 16;
            unsigned int w[kTrDigits][4];
#pragma unroll
            for (int i = 0; i < kTrDigits; ++i) w[i][0] = w[i][1] = w[i][2] = w[i][3] = 0u;
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                const long long r = r0 + tt;
                long long q = 0;
                if (live && r < n) q = __double2ll_rn((x[r * row_stride + col] - m) * sc);
#pragma unroll
                for (int i = 0; i < kTrDigits; ++i) {
                    const int d = (int)(signed char)(q & 0xFF);                 // balanced digit in [-128, 127]
                    q = (q - d) >> 8;
                    w[i][tt >> 2] |= (unsigned int)(d & 0xFF) << (8 * (tt & 3));
                }
            }
#pragma unroll
            for (int i = 0; i < kTrDigits; ++i)
                *reinterpret_cast<uint4*>(digits + ((size_t)i * n_groups + (g0 + g)) * kTrLboD + (size_t)t * 16) =
                    make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
        }
    } else if (t - kTrF < n_bins) {
        const int b = t - kTrF;
        for (int g = 0; g < kTrGroups; ++g) {
            const long long r0 = (g0 + g) * 16;
            int lab[16];
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                const long long r = r0 + tt;
                int k = -1;
                if (r < n) {
                    k = (int)labels[r * n_bins + b];
                    k = k < 0 ? 0 : (k >= n_classes ? n_classes - 1 : k);
                    atomicAdd(&s_cnt[b * n_classes + k], 1u);
                }
                lab[tt] = k;
            }
            for (int k = 0; k < n_classes; ++k) {
                unsigned int w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int tt = 0; tt < 16; ++tt) w[tt >> 2] |= (lab[tt] == k ? 1u : 0u) << (8 * (tt & 3));
                *reinterpret_cast<uint4*>(onehot + (size_t)(g0 + g) * kTrLboO + (size_t)(b * n_classes + k) * 16) =
                    make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    __syncthreads();
    for (int i = t; i < n_bins * n_classes; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(counts + i, (unsigned long long)s_cnt[i]);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

// grid = (splits, 27 roles); block = 128 threads
__global__ void __launch_bounds__(128, 1)
k_tc_stats(const unsigned char* __restrict__ digits, const unsigned char* __restrict__ onehot, long long n_groups, int n_chunks,
           int chunks_per_split, long long* __restrict__ acc_out /*[role][3][128][160]*/) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* tail = smem + kTrStages * kTrStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);                 // full[3], empty[3], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kTrStages]), bar_done = smem_u32(&bars[2 * kTrStages]);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = blockIdx.y;
    const int c0 = blockIdx.x * chunks_per_split, c1 = min(n_chunks, c0 + chunks_per_split);
    if (c0 >= c1) return;
    const bool gram = role < kTrPairs;
    int di = 0, dj = 0;
    if (gram) tr_pair(role, di, dj); else dj = role - kTrPairs;
    const int n_acc = gram ? 2 : 3;
    // stage layout: [A operand | B operand]; Gram with i == j loads one chunk and uses it for both
    const uint32_t a_bytes = gram ? kTrDChunk : kTrOChunk;
    const bool shared_ab = gram && di == dj;
    const uint32_t b_off = shared_ab ? 0u : a_bytes;
    const uint32_t stage_tx = shared_ab ? (uint32_t)kTrDChunk : a_bytes + kTrDChunk;

    if (tid == 0) {
        for (int i = 0; i < kTrStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        const unsigned char* srcA = gram ? digits + (size_t)di * n_groups * kTrLboD : onehot;
        const size_t strideA = gram ? kTrDChunk : kTrOChunk;
        const unsigned char* srcB = digits + (size_t)dj * n_groups * kTrLboD;
        uint32_t it = 0;
        for (int c = c0; c < c1; ++c, ++it) {
            const uint32_t s = it % kTrStages, n_use = it / kTrStages;
            if (n_use > 0) mbar_wait(bar_empty + 8 * s, (n_use - 1) & 1);
            mbar_expect_tx(bar_full + 8 * s, stage_tx);
            const uint32_t dst = smem_u32(smem + s * kTrStage);
            tma_bulk_load(dst, srcA + (size_t)c * strideA, a_bytes, bar_full + 8 * s);
            if (!shared_ab) tma_bulk_load(dst + b_off, srcB + (size_t)c * kTrDChunk, kTrDChunk, bar_full + 8 * s);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        // instruction descriptor: D s32, A/B s8, both K-major, N = 160, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTrF >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t lbo_a = gram ? kTrLboD : kTrLboO;
        uint32_t it = 0;
        for (int c = c0; c < c1; ++c, ++it) {
            const uint32_t s = it % kTrStages;
            mbar_wait(bar_full + 8 * s, (it / kTrStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + s * kTrStage), sb = sa + b_off;
#pragma unroll
            for (int ks = 0; ks < kTrChunkRows / 32; ++ks) {                    // K = 32 rows = two 16-row groups
                const uint64_t db = umma_desc_kmajor(sb + ks * 2 * kTrLboD, kTrLboD, kTrSbo);
                for (int a = 0; a < n_acc; ++a) {
                    // M tile a: Gram -> feature columns [0,128) / [32,160); class sums -> one-hot columns [128a, 128a+128)
                    const uint32_t col0 = gram ? (a ? 32u : 0u) : 128u * a;
                    const uint64_t da = umma_desc_kmajor(sa + ks * 2 * lbo_a + col0 * 16, lbo_a, kTrSbo);
                    mma_i8(tmem + a * kTrF, da, db, idesc, (it == 0 && ks == 0) ? 0u : 1u);
                }
            }
            umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_done);
    }
    // ===== epilogue: every thread owns accumulator row `tid` =====
    __syncwarp();
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int a = 0; a < n_acc; ++a) {
        unsigned long long* out = reinterpret_cast<unsigned long long*>(acc_out) + ((size_t)role * kTrAccPerRole + a) * kTrAccElems +
                                  (size_t)tid * kTrF;
#pragma unroll 1
        for (int cc = 0; cc < kTrF; cc += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + a * kTrF + cc, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (v[j]) atomicAdd(out + cc + j, (unsigned long long)(long long)(int)v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// ---- integer accumulators -> fp64 statistics ---------------------------------------------------------------------------
__device__ __forceinline__ long long tr_gram_acc(const long long* acc, int role, int p, int q) {
    const int a = p < 128 ? 0 : 1, row = p - (a ? 32 : 0);
    return acc[((size_t)role * kTrAccPerRole + a) * kTrAccElems + (size_t)row * kTrF + q];
}

__global__ void k_tc_combine(const long long* __restrict__ acc, const unsigned long long* __restrict__ counts_in,
                             const double* __restrict__ from_fixed, int nf, int n_bins, int n_classes, double* __restrict__ G,
                             double* __restrict__ sums, double* __restrict__ counts) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_g = (long long)nf * nf, n_s = (long long)n_bins * n_classes * nf;
    if (idx < n_g) {
        const int p = (int)(idx / nf), q = (int)(idx - (long long)p * nf);
        long long T[2 * kTrDigits - 1];
#pragma unroll
        for (int c = 0; c < 2 * kTrDigits - 1; ++c) T[c] = 0;
        int role = 0;
        for (int i = 0; i < kTrDigits; ++i)
            for (int j = i; j < kTrDigits; ++j, ++role) {
                long long t = tr_gram_acc(acc, role, p, q);
                if (i != j) t += tr_gram_acc(acc, role, q, p);          // D_j^T D_i = (D_i^T D_j)^T
                T[i + j] += t;
            }
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 2 * kTrDigits - 1; ++c) s += (double)T[c] * ldexp(1.0, 8 * c);     // ascending magnitude
        G[idx] = s * from_fixed[p] * from_fixed[q];
    } else if (idx < n_g + n_s) {
        const long long k = idx - n_g;
        const int bk = (int)(k / nf), f = (int)(k - (long long)bk * nf);
        double s = 0.0;
        for (int j = 0; j < kTrDigits; ++j) {
            const long long t = acc[((size_t)(kTrPairs + j) * kTrAccPerRole + bk / 128) * kTrAccElems + (size_t)(bk % 128) * kTrF + f];
            s += (double)t * ldexp(1.0, 8 * j);
        }
        sums[k] = s * from_fixed[f];
    } else if (idx < n_g + n_s + (long long)n_bins * n_classes) {
        const long long k = idx - n_g - n_s;
        counts[k] = (double)counts_in[k];
    }
}

bool lda_stats_tc_supported(int nf, int n_bins, int n_classes) {
    return nf <= kTrF && n_bins * n_classes <= kTrOH && n_bins <= 256 - kTrF;
}

// Same contract as lda_stats_run (train.cu); xbar must already hold the centre.
int lda_stats_tc_run(const double* x, long long n, long long row_stride, const int* select, int nf, const double* labels, int n_bins,
                     int n_classes, const double* xbar, double* G, double* sums, double* counts, cudaStream_t st) {
    const int n_chunks = ceil_div(n, kTrChunkRows);
    const long long n_groups = (long long)n_chunks * kTrGroups;
    int splits = ceil_div(n_chunks, kTrMaxChunksPerSplit);
    if (splits < 5) splits = n_chunks < 5 ? n_chunks : 5;               // 27 roles x 5 splits = 135 CTAs on 148 SMs
    const int chunks_per_split = ceil_div(n_chunks, splits);
    unsigned char *digits = nullptr, *onehot = nullptr;
    long long* acc = nullptr;
    unsigned long long *amax = nullptr, *cnt = nullptr;
    double *to_fixed = nullptr, *from_fixed = nullptr;
    const size_t acc_bytes = sizeof(long long) * kTrRoles * kTrAccPerRole * kTrAccElems;
    SGS_CUDA(cudaMallocAsync((void**)&digits, (size_t)kTrDigits * n_groups * kTrLboD, st));
    SGS_CUDA(cudaMallocAsync((void**)&onehot, (size_t)n_groups * kTrLboO, st));
    SGS_CUDA(cudaMallocAsync((void**)&acc, acc_bytes, st));
    SGS_CUDA(cudaMallocAsync((void**)&amax, sizeof(unsigned long long) * kTrF, st));
    SGS_CUDA(cudaMallocAsync((void**)&cnt, sizeof(unsigned long long) * kTrOH, st));
    SGS_CUDA(cudaMallocAsync((void**)&to_fixed, sizeof(double) * kTrF, st));
    SGS_CUDA(cudaMallocAsync((void**)&from_fixed, sizeof(double) * kTrF, st));
    SGS_CUDA(cudaMemsetAsync(acc, 0, acc_bytes, st));
    SGS_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned long long) * kTrF, st));
    SGS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * kTrOH, st));
    SGS_CUDA(cudaMemsetAsync(onehot, 0, (size_t)n_groups * kTrLboO, st));       // columns past bins * classes stay zero
    static unsigned long long optin = 0;
    SGS_CUDA(smem_optin(k_tc_stats, kTrSmem, &optin));
    int slices = (int)(n / 4096);
    slices = slices < 1 ? 1 : (slices > 592 ? 592 : slices);
    k_tc_absmax<<<slices, kTrF, 0, st>>>(x, select, xbar, n, row_stride, nf, amax);
    SGS_LAUNCHED();
    k_tc_scales<<<1, kTrF, 0, st>>>(amax, nf, to_fixed, from_fixed);
    SGS_LAUNCHED();
    k_tc_pack<<<n_chunks, 256, sizeof(unsigned int) * n_bins * n_classes, st>>>(x, select, xbar, to_fixed, labels, n, row_stride, nf, n_bins,
                                                                                 n_classes, n_groups, digits, onehot, cnt);
    SGS_LAUNCHED();
    {
        ProfScope ps(kProfTrainTc, st);
        k_tc_stats<<<dim3(splits, kTrRoles), 128, kTrSmem, st>>>(digits, onehot, n_groups, n_chunks, chunks_per_split, acc);
    }
    SGS_LAUNCHED();
    const long long n_out = (long long)nf * nf + (long long)n_bins * n_classes * (nf + 1);
    k_tc_combine<<<ceil_div(n_out, 256), 256, 0, st>>>(acc, cnt, from_fixed, nf, n_bins, n_classes, G, sums, counts);
    SGS_LAUNCHED();
    cudaFreeAsync(digits, st); cudaFreeAsync(onehot, st); cudaFreeAsync(acc, st); cudaFreeAsync(amax, st); cudaFreeAsync(cnt, st);
    cudaFreeAsync(to_fixed, st); cudaFreeAsync(from_fixed, st);
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
