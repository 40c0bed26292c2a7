// LDA class scoring on the 5th-generation tensor cores (tcgen05, accumulators in TMEM, operands by TMA bulk copy).
//
//   scores[frame][bin*9 + class] = (x[frame] - c) . W'[bin, class]        (bias' = bias + W c is added in the epilogue)
//
// is a dense (frames x 160) . (160 x 384) contraction.  tcgen05 has no fp64 kind, so the product runs as split TF32:
// x' = x_hi + x_lo, W' = W_hi + W_lo (each part a properly rounded tf32), accumulated as hi.hi + hi.lo + lo.hi in the
// fp32 TMEM accumulator (24 bits of each operand).  That is a FILTER, not the answer: the epilogue takes the per-bin
// argmax and flags every frame in which the best and second-best class of some bin are closer than a bound on the
// tensor-core error; the flagged (frame, bin) pairs - a bit mask per frame and slice - are re-scored exactly in fp64
// (lda.cu:k_lda_pairs).  Class indices therefore equal the fp64 result everywhere (LDASynthesis.py:25-26 semantics), not
// just "up to near-ties".  With a TRAINED model 13 % of the frames have a near-tie in some bin (random weights: 0.003 %),
// but only ~1 of their 40 bins: re-scoring pairs instead of whole frames is what keeps the exact pass off the step time.
//
// k_lda_pack  full-occupancy pre-pass: gathers the stacked, selected features of every 128-frame tile straight from
//             the un-stacked log-power array, centres them, splits hi/lo and writes them to HBM already in the
//             canonical K-major core-matrix layout the MMA reads ([tile][hi|lo][K/4][128 rows x 16 B]); also |x'|^2.
// k_lda_tc    persistent, warp-specialised.  The 360 outputs are cut into 3 slices of <= 14 bins (N = 128); a CTA owns
//             one slice and keeps its W' slice (hi + lo, 160 KB) in shared memory for its whole life.
//               warp 4, one lane : TMA producer - cp.async.bulk of the next K = 16 chunk (8 KB hi + 8 KB lo) into a
//                                  4-stage ring, completion counted on the stage's "full" mbarrier
//               warp 5, one lane : MMA issuer - 6 tcgen05.mma (M128 N128 K8, kind::tf32) per chunk, tcgen05.commit to the
//                                  stage's "empty" mbarrier; accumulators double-buffered in TMEM (2 x 128 columns)
//               warps 0-3        : epilogue - tcgen05.ld of the finished accumulator (thread = frame row), bias, per-bin
//                                  argmax, near-tie flag, label store; overlaps the next tile's loads and MMAs
#include <math.h>
#include "kernels.cuh"
#include "tc.cuh"

namespace sgs {

constexpr int kTcM = 128, kTcN = 128, kTcK = 160, kTcChunk = 16, kTcStages = 4, kTcClasses = 9;
constexpr int kTcThreads = 192;
constexpr uint32_t kSBO = 128;                          // bytes between 8-row groups of core matrices
constexpr uint32_t kLBO = (kTcM / 8) * 128;             // bytes between consecutive 16-byte K groups (2048)
constexpr int kBBytes = (kTcK / 4) * kLBO;              // one 128 x 160 tf32 operand matrix: 81 920 B
constexpr int kHalfStage = (kTcChunk / 4) * kLBO;       // hi (or lo) part of one chunk: 8 192 B
constexpr int kAStageBytes = 2 * kHalfStage;
constexpr int kTileBytes = 2 * kBBytes;                 // packed A tile in HBM: hi matrix then lo matrix
constexpr int kTcTail = 256 + kTcN * 8 + 16 * 8 + kTcN * 4 + 2 * 16 * 4;   // barriers + bias + per-bin weight norms + fp32 copies
constexpr int kTcSmem = 2 * kBBytes + kTcStages * kAStageBytes + kTcTail;


__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    // K-major, no swizzle: start address, leading (K) byte offset, stride (row-group) byte offset, version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((kLBO >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((kSBO >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

// ---- pre-pass: packed, centred, split operand tiles + row norms ------------------------------------------------
// block = 512 threads = 128 rows x 4 K-groups in flight; grid = tiles
__global__ void __launch_bounds__(512)
k_lda_pack(const double* __restrict__ feat, const int* __restrict__ feat_chan, const int* __restrict__ feat_back,
           const double* __restrict__ centre, float* __restrict__ packed, double* __restrict__ xnorm2, const LdaTcGeom g) {
    __shared__ double s_norm[4][kTcM];
    const int tile = blockIdx.x;
    const int sess = tile / g.tiles_per_session;
    const int r = threadIdx.x & (kTcM - 1), q = threadIdx.x >> 7;
    const int row = (tile - sess * g.tiles_per_session) * kTcM + r;
    const bool live = row < g.n_rows;
    const double* fs = feat + (long long)sess * g.n_windows * g.n_channels;
    unsigned char* out = reinterpret_cast<unsigned char*>(packed) + (size_t)tile * kTileBytes;
    const uint32_t row_off = (uint32_t)(r / 8) * kSBO + (uint32_t)(r % 8) * 16;
    double n2 = 0.0;
    for (int kg = q; kg < kTcK / 4; kg += 4) {
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = kg * 4 + j;
            double v = 0.0;
            if (live && f < g.n_features) {
                const int w = row + g.first_row - __ldg(feat_back + f);
                v = (w >= 0 ? __ldg(fs + (long long)w * g.n_channels + __ldg(feat_chan + f)) : 0.0) - __ldg(centre + f);
            }
            n2 = fma(v, v, n2);
            hi[j] = to_tf32((float)v);
            lo[j] = to_tf32((float)(v - (double)hi[j]));
        }
        *reinterpret_cast<float4*>(out + (size_t)kg * kLBO + row_off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(out + kBBytes + (size_t)kg * kLBO + row_off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    s_norm[q][r] = n2;
    __syncthreads();
    if (q == 0) xnorm2[(size_t)tile * kTcM + r] = (s_norm[0][r] + s_norm[1][r]) + (s_norm[2][r] + s_norm[3][r]);
}

// grid = (ctas_per_slice, 3); block = 192 threads: warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
__global__ void __launch_bounds__(kTcThreads, 1)
k_lda_tc(const float* __restrict__ packed, const double* __restrict__ xnorm2, const float* __restrict__ Bmat /*[3][hi,lo][canonical]*/,
         const double* __restrict__ bias /*[3][128] incl. centring, -inf padding*/, const double* __restrict__ cls /*[3][128]*/,
         const int* __restrict__ slice_bins /*[4]*/, const double* __restrict__ wnorm /*[3][16]*/, double* __restrict__ labels,
         int* __restrict__ flags, const LdaTcGeom g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;                                           // hi then lo
    unsigned char* sA = smem + 2 * kBBytes;                             // stages x (hi, lo)
    unsigned char* tail = sA + kTcStages * kAStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);                 // full[4], empty[4], acc_full[2], acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
    double* s_bias = reinterpret_cast<double*>(tail + 256);             // [128]
    double* s_wnorm = s_bias + kTcN;                                    // [16]
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kTcStages]);
    const uint32_t bar_acc_full = smem_u32(&bars[2 * kTcStages]), bar_acc_empty = smem_u32(&bars[2 * kTcStages + 2]);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slice = blockIdx.y;
    const int bin0 = slice_bins[slice], nb = slice_bins[slice + 1] - bin0;

    // ---- one-time setup ------------------------------------------------------------------------------------
    {
        const float4* src = reinterpret_cast<const float4*>(Bmat + (size_t)slice * 2 * (kBBytes / 4));
        float4* dst = reinterpret_cast<float4*>(sB);
        for (int i = tid; i < 2 * kBBytes / 16; i += blockDim.x) dst[i] = src[i];
        for (int i = tid; i < kTcN; i += blockDim.x) s_bias[i] = bias[slice * kTcN + i];
        if (tid < 16) s_wnorm[tid] = wnorm[slice * 16 + tid];
    }
    if (tid == 0) {
        for (int i = 0; i < kTcStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * kTcN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // W' slice (generic-proxy stores) -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    constexpr int kChunks = kTcK / kTcChunk;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(packed) + (size_t)tile * kTileBytes;
                for (int c = 0; c < kChunks; ++c, ++it) {
                    const uint32_t s = it % kTcStages, n_use = it / kTcStages;
                    if (n_use > 0) mbar_wait(bar_empty + 8 * s, (n_use - 1) & 1);       // MMAs of the previous use are done
                    mbar_expect_tx(bar_full + 8 * s, kAStageBytes);
                    const uint32_t dst = smem_u32(sA + s * kAStageBytes);
                    tma_bulk_load(dst, src + (size_t)c * kHalfStage, kHalfStage, bar_full + 8 * s);
                    tma_bulk_load(dst + kHalfStage, src + kBBytes + (size_t)c * kHalfStage, kHalfStage, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
            const uint32_t b_hi = smem_u32(sB), b_lo = b_hi + kBBytes;
            uint32_t it = 0, t_local = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++t_local) {
                const uint32_t buf = t_local & 1, n_acc = t_local >> 1;
                if (n_acc > 0) mbar_wait(bar_acc_empty + 8 * buf, (n_acc - 1) & 1);     // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem + buf * kTcN;
                for (int c = 0; c < kChunks; ++c, ++it) {
                    const uint32_t s = it % kTcStages;
                    mbar_wait(bar_full + 8 * s, (it / kTcStages) & 1);                  // chunk landed in shared memory
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(sA + s * kAStageBytes), a_lo = a_hi + kHalfStage;
#pragma unroll
                    for (int ks = 0; ks < kTcChunk / 8; ++ks) {
                        const uint32_t a_off = ks * 2 * kLBO, b_off = (uint32_t)(c * (kTcChunk / 8) + ks) * 2 * kLBO;
                        const uint64_t dah = umma_desc(a_hi + a_off), dal = umma_desc(a_lo + a_off);
                        const uint64_t dbh = umma_desc(b_hi + b_off), dbl = umma_desc(b_lo + b_off);
                        mma_tf32(acc, dal, dbh, idesc, (c == 0 && ks == 0) ? 0u : 1u);  // small cross terms first, then hi.hi
                        mma_tf32(acc, dah, dbl, idesc, 1u);
                        mma_tf32(acc, dah, dbh, idesc, 1u);
                    }
                    umma_commit(bar_empty + 8 * s);                                     // stage reusable when these MMAs finish
                }
                umma_commit(bar_acc_full + 8 * buf);                                    // accumulator complete -> epilogue
            }
        }
    } else {
        // ===== epilogue (warps 0-3): thread = accumulator row, warp w owns TMEM lanes 32w.. =====
        // Fully unrolled over the 126 score columns (bin = column / 9, class = column % 9 are compile-time), comparisons in fp32:
        // the first version walked the columns with run-time bin / class counters in fp64 - 3650 instructions per thread and tile,
        // branch_resolving the second largest stall, and with one epilogue warp per scheduler that was the kernel's critical path
        // (tensor pipe 17 % busy).  The extra rounding of the fp32 bias add is part of the tie threshold below.
        float* s_biasf = reinterpret_cast<float*>(s_wnorm + 16);            // [128] bias' as float
        float* s_tol = s_biasf + kTcN;                                      // [16]  per-bin 2^-22 max|bias'| (rounding of the bias)
        float* s_wnf = s_tol + 16;                                          // [16]  per-bin weight norm, rounded up
        for (int i = tid; i < kTcN; i += 128) s_biasf[i] = (float)s_bias[i];
        if (tid < 16) {
            double bm = 0.0;
            for (int k = 0; k < kTcClasses; ++k) {
                const double bv = tid * kTcClasses + k < kTcN ? s_bias[tid * kTcClasses + k] : 0.0;
                if (isfinite(bv)) bm = fmax(bm, fabs(bv));
            }
            s_tol[tid] = (float)(bm * 2.4e-7) * 1.0001f;
            s_wnf[tid] = (float)s_wnorm[tid] * 1.0001f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");                      // epilogue warps only
        uint32_t t_local = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++t_local) {
            const uint32_t buf = t_local & 1;
            const int sess = tile / g.tiles_per_session;
            const int erow = (tile - sess * g.tiles_per_session) * kTcM + tid;
            const bool elive = erow < g.n_rows;
            // x |w|_2 of the bin bounds sum |x' w|; rounded up when narrowed to float
            const float margin0 = (float)(2.0 * g.eps * sqrt(xnorm2[(size_t)tile * kTcM + tid])) * 1.0001f;
            double* lab_row = labels + ((long long)sess * g.n_rows + erow) * g.n_bins + bin0;
            const double* cls_s = cls + slice * kTcN;
            mbar_wait(bar_acc_full + 8 * buf, (t_local >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned tiemask = 0;                                       // bins of this slice whose argmax the filter cannot decide
            float best = -INFINITY, second = -INFINITY;
            int best_k = 0;
#pragma unroll
            for (int c0 = 0; c0 < kTcN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + buf * kTcN + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    constexpr int kBinsPerSlice = kTcN / kTcClasses;        // 14
                    const int col = c0 + j, bin = col / kTcClasses, kk = col - bin * kTcClasses;
                    if (bin < kBinsPerSlice) {
                        const float sc = __uint_as_float(v[j]) + s_biasf[col];
                        const bool gt = sc > best;                          // strict: the first maximum wins, as numpy argmax
                        second = gt ? best : fmaxf(second, sc);
                        best_k = gt ? kk : best_k;
                        best = gt ? sc : best;
                        if (kk == kTcClasses - 1) {
                            if (bin < nb) {
                                // |fp32 score - exact| <= eps |x'| |w| (tensor-core split) + rounding of the bias and of the fp32 add
                                const float fin2 = second > -INFINITY ? fabsf(second) : 0.0f;
                                const float thr = fmaf(margin0, s_wnf[bin], s_tol[bin]) + 4.8e-7f * (fabsf(best) + fin2);
                                if (best - second < thr) tiemask |= 1u << bin;
                                if (elive) lab_row[bin] = cls_s[bin * kTcClasses + best_k];
                            }
                            best = -INFINITY; second = -INFINITY; best_k = 0;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_acc_empty + 8 * buf);                       // 128 arrivals free the accumulator
            if (elive && tiemask) flags[((long long)sess * g.n_rows + erow) * 3 + slice] = (int)tiemask;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * kTcN));
}

// per call: centring vector c[f] = mean log-power of the feature's channel (any constant vector is valid: it is folded
// into the bias), bias'[slice][n] = bias + W . c in fp64, -inf on padding rows
__global__ void k_lda_tc_prep(const double* __restrict__ chan_mean, const int* __restrict__ feat_chan, const double* __restrict__ Wt /*[bin][F][9]*/,
                              const double* __restrict__ bias0 /*[bin][9]*/, const int* __restrict__ slice_bins, int n_features,
                              double* __restrict__ centre /*[160]*/, double* __restrict__ bias_out /*[3][128]*/) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < kTcK) centre[t] = t < n_features ? chan_mean[feat_chan[t]] : 0.0;
    if (t < 3 * kTcN) {
        const int slice = t / kTcN, n = t - slice * kTcN;
        const int bin = slice_bins[slice] + n / kTcClasses, k = n % kTcClasses;
        double b = -INFINITY;
        if (bin < slice_bins[slice + 1]) {
            b = bias0[bin * kTcClasses + k];
            const double* w = Wt + (long long)bin * n_features * kTcClasses + k;
            for (int f = 0; f < n_features; ++f) b = fma(w[f * kTcClasses], chan_mean[feat_chan[f]], b);
        }
        bias_out[t] = b;
    }
}

// compact the flagged (frame, slice) entries into a list (order does not matter); count[0] = entries, count[1] = flagged (frame, bin) pairs
__global__ void k_flag_list(const int* __restrict__ flags, long long n, int* __restrict__ list, int* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = i < n ? flags[i] : 0;
    if (m) {
        list[atomicAdd(count, 1)] = (int)i;
        atomicAdd(count + 1, __popc((unsigned)m));
    }
}

int lda_tc_run(const double* feat, const float* Bmat, const double* Wt, const double* bias0, const double* chan_mean, double* bias,
               const double* cls, const int* feat_chan, const int* feat_back, double* centre, const int* slice_bins, const double* wnorm,
               double* labels, int* flags, int* list, int* count, long long n_frames_total, const LdaTcGeom& g, cudaStream_t st) {
    float* packed = nullptr;
    double* xnorm2 = nullptr;
    SGS_CUDA(cudaMallocAsync((void**)&packed, (size_t)g.n_tiles * kTileBytes, st));
    SGS_CUDA(cudaMallocAsync((void**)&xnorm2, sizeof(double) * (size_t)g.n_tiles * kTcM, st));
    static unsigned long long optin = 0;
    SGS_CUDA(smem_optin(k_lda_tc, kTcSmem, &optin));
    SGS_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 3 * n_frames_total, st));
    SGS_CUDA(cudaMemsetAsync(count, 0, 2 * sizeof(int), st));
    k_lda_tc_prep<<<ceil_div(3 * kTcN, 128), 128, 0, st>>>(chan_mean, feat_chan, Wt, bias0, slice_bins, g.n_features, centre, bias);
    SGS_LAUNCHED();
    const int per_slice = g.n_tiles < 49 ? g.n_tiles : 49;              // 3 x 49 = 147 persistent CTAs on 148 SMs
    {
        ProfScope ps(kProfLdaPack, st);
        k_lda_pack<<<g.n_tiles, 512, 0, st>>>(feat, feat_chan, feat_back, centre, packed, xnorm2, g);
    }
    SGS_LAUNCHED();
    {
        ProfScope ps(kProfLdaTc, st);
        k_lda_tc<<<dim3(per_slice, 3), kTcThreads, kTcSmem, st>>>(packed, xnorm2, Bmat, bias, cls, slice_bins, wnorm, labels, flags, g);
    }
    SGS_LAUNCHED();
    cudaFreeAsync(packed, st);
    cudaFreeAsync(xnorm2, st);
    k_flag_list<<<ceil_div(3 * n_frames_total, 256), 256, 0, st>>>(flags, 3 * n_frames_total, list, count);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
