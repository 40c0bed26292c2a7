// Probe: one tcgen05 kind::tf32 tile D[128 x 128] = A[128 x K] * B[128 x K]^T, operands written to shared memory by
// ordinary threads in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices), accumulator in TMEM,
// read back with tcgen05.ld.  Validates descriptor encodings before the real LDA kernel is built on them.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int M = 128, N = 128, K = 32;          // K = 4 MMA k-steps of 8
constexpr uint32_t SBO = 128;                    // bytes between 8-row groups
constexpr uint32_t LBO = (M / 8) * 128;          // bytes between the two 16-byte K halves of one k-step

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);              // start address, bits [0,14)
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;          // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;          // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                              // version = 1 (Blackwell)
    return d;                                            // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 2u << 7;                    // a_format = TF32
    d |= 2u << 10;                   // b_format = TF32
    d |= (uint32_t)(n >> 3) << 17;   // n_dim
    d |= (uint32_t)(m >> 4) << 24;   // m_dim
    return d;                        // a_major = b_major = K (0), no negate, dense
}

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                         // M x K canonical
    float* sB = reinterpret_cast<float*>(smem + M * K * 4);             // N x K canonical
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < M * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        const uint32_t off = (k / 4) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 4) * 4;
        sA[off / 4] = A[i];
        sB[off / 4] = B[i];                                             // N == M here
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");                     // generic-proxy smem writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, N);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO, LBO, SBO);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO, LBO, SBO);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    // everyone waits for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // warp w reads TMEM lanes 32w..32w+31 (row = 32w + lane), 128 fp32 columns in 4 chunks of 32
    const int row = tid;
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
        for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

int main() {
    float *hA = new float[M * K], *hB = new float[N * K], *hD = new float[M * N];
    srand(1);
    auto tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    for (int i = 0; i < M * K; ++i) hA[i] = tf32((rand() % 2001 - 1000) / 500.0f);
    for (int i = 0; i < N * K; ++i) hB[i] = tf32((rand() % 2001 - 1000) / 500.0f);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4);
    const int smem = (M + N) * K * 4;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)hA[i * K + k] * hB[j * K + k];
            maxerr = fmax(maxerr, fabs(ref - hD[i * N + j])); maxref = fmax(maxref, fabs(ref));
        }
    printf("max |err| = %.3g (max |ref| = %.3g)  D[0][0..3] = %g %g %g %g\n", maxerr, maxref, hD[0], hD[1], hD[2], hD[3]);
    printf(maxerr < 1e-4 * maxref ? "PROBE OK\n" : "PROBE MISMATCH\n");
    return 0;
}
