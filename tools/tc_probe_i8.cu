// Probe: tcgen05 kind::i8 (s8 x s8 -> s32) on the operand layout the LDA-training Gram kernel uses.
// Packed operand P[kgroup][160 columns][16 bytes]: for every group of 16 consecutive rows (the contraction index) the 16
// int8 values of one column are contiguous - the canonical K-major no-swizzle core-matrix layout with SBO = 128 B (8
// columns) and LBO = 160 * 16 B (next 16 rows).  D[128 x 160] = A^T B with A = columns [32, 160), B = columns [0, 160):
// one buffer serves as both operands, A merely starts 32 columns in.  K = 128 rows = 4 MMAs of K = 32.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe_i8 tc_probe_i8.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

constexpr int COLS = 160, M = 128, N = 160, K = 128, A0 = 32;
constexpr uint32_t SBO = 128, LBO = COLS * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) | ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128) k_probe(const int8_t* __restrict__ P, int* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (K / 16) * COLS * 16; i += blockDim.x) smem[i] = (unsigned char)P[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        // c_format S32 (2), a/b format INT8 (1), K-major both, N, M
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < K / 32; ++ks) {
            const uint32_t base = smem_u32(smem) + ks * 2 * LBO;            // 32 rows = 2 groups of 16
            const uint64_t da = make_desc(base + A0 * 16), db = make_desc(base);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0));
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

int main() {
    static int8_t X[K][COLS], P[(K / 16) * COLS * 16];
    static int D[M * N];
    srand(3);
    for (int t = 0; t < K; ++t)
        for (int c = 0; c < COLS; ++c) {
            X[t][c] = (int8_t)(rand() % 256 - 128);
            P[((t / 16) * COLS + c) * 16 + t % 16] = X[t][c];
        }
    int8_t* dP; int* dD;
    cudaMalloc(&dP, sizeof(P)); cudaMalloc(&dD, sizeof(D));
    cudaMemcpy(dP, P, sizeof(P), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof(D));
    const int smem = sizeof(P);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 128, smem>>>(dP, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(D, dD, sizeof(D), cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            int ref = 0;
            for (int t = 0; t < K; ++t) ref += (int)X[t][A0 + i] * (int)X[t][j];
            if (ref != D[i * N + j]) { if (bad < 5) printf("D[%d][%d] = %d, want %d\n", i, j, D[i * N + j], ref); ++bad; }
        }
    printf("%lld of %d entries differ\n", bad, M * N);
    printf(bad == 0 ? "PROBE OK (s8 x s8 -> s32 exact)\n" : "PROBE MISMATCH\n");
    return 0;
}
