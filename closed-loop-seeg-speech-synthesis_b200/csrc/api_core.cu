// Library plumbing: error reporting, host/device buffer staging, scratch memory, init.
#include <stdarg.h>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "../../include/sgs.h"

namespace sgs {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
bool g_prof_on = false;
struct ProfSpan { int id; cudaEvent_t a, b; };
static std::vector<ProfSpan> g_spans;
static cudaEvent_t g_open[kProfCount];
static double g_prof_ms[kProfCount];
static unsigned long long g_prof_n[kProfCount];
static const char* kProfNames[kProfCount] = {"iir_init", "iir_state", "iir_carry", "iir_feat", "stack", "lda", "gl_blocks", "gl_ola",
                                             "lowpass", "stream", "gl_batch", "logmel", "train", "lda_tc", "train_tc", "iir_pieces_state", "iir_pieces_feat", "lda_pack", "iir_pieces_tail"};

void prof_begin(int id, cudaStream_t st) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    g_open[id] = e;
}
void prof_end(int id, cudaStream_t st) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    g_spans.push_back({id, g_open[id], e});
}
static void prof_collect() {
    for (auto& s : g_spans) {
        cudaEventSynchronize(s.b);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { g_prof_ms[s.id] += ms; g_prof_n[s.id] += 1; }
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    g_spans.clear();
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return SGS_ERR_CUDA;
}

__global__ void k_copy_in16(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) dst[i] = src[i];
}

int copy_in_small(void* dst, const void* src, size_t bytes, bool src_pinned, cudaStream_t st) {
    if (bytes == 0) return SGS_OK;
    if (src_pinned && bytes % 16 == 0 && ((uintptr_t)dst | (uintptr_t)src) % 16 == 0 && bytes <= (1u << 20)) {
        const size_t n16 = bytes / 16;
        k_copy_in16<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((uint4*)dst, (const uint4*)src, n16);   // UVA: the pinned host pointer is valid on the device
        SGS_LAUNCHED();
        SGS_CUDA(cudaGetLastError());
        return SGS_OK;
    }
    SGS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    return SGS_OK;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int stage_in(Staged& s, const void* p, size_t bytes, cudaStream_t st) {
    s = Staged();
    s.bytes = bytes;
    if (bytes == 0) return SGS_OK;
    if (is_device_ptr(p)) { s.dev = const_cast<void*>(p); return SGS_OK; }
    s.host = const_cast<void*>(p);
    SGS_CUDA(cudaMallocAsync(&s.dev, bytes, st));
    s.owned = true;
    SGS_CUDA(cudaMemcpyAsync(s.dev, p, bytes, cudaMemcpyHostToDevice, st));
    return SGS_OK;
}

int stage_out(Staged& s, void* p, size_t bytes, cudaStream_t st) {
    s = Staged();
    s.bytes = bytes;
    if (bytes == 0) return SGS_OK;
    if (is_device_ptr(p)) { s.dev = p; return SGS_OK; }
    s.host = p;
    SGS_CUDA(cudaMallocAsync(&s.dev, bytes, st));
    s.owned = true;
    return SGS_OK;
}

int finish_out(Staged& s, cudaStream_t st) {
    if (s.host && s.bytes) SGS_CUDA(cudaMemcpyAsync(s.host, s.dev, s.bytes, cudaMemcpyDeviceToHost, st));
    return SGS_OK;
}

void release(Staged& s, cudaStream_t st) {
    if (s.owned && s.dev) cudaFreeAsync(s.dev, st);
    s = Staged();
}

}  // namespace sgs

extern "C" {

int sgs_abi_version(void) { return SGS_ABI_VERSION; }
const char* sgs_last_error(void) { return sgs::g_err; }
unsigned long long sgs_launch_count(void) { return sgs::g_launches; }

int sgs_device_count(int* count) {
    SGS_ARG(count != nullptr, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return sgs::cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    return SGS_OK;
}

int sgs_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        sgs::set_error("no CUDA device available (libsgs has no CPU fallback)");
        return SGS_ERR_CUDA;
    }
    SGS_ARG(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    SGS_CUDA(cudaSetDevice(device));
    SGS_CUDA(cudaFree(0));
    cudaDeviceProp p;
    SGS_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        sgs::set_error("libsgs is built for sm_100a only; device %d is sm_%d%d", device, p.major, p.minor);
        return SGS_ERR_UNSUPPORTED;
    }
    // keep freed stream-ordered scratch around instead of returning it to the OS after every call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return SGS_OK;
}

/* Kernel-class timing: enable, run, then read.  name in {iir_init, iir_state, iir_carry, iir_feat, stack, lda, gl_blocks,
 * gl_ola, lowpass, stream, gl_batch, logmel, train, lda_tc, train_tc, iir_pieces_state, iir_pieces_feat, lda_pack}.
 * iir_state / iir_feat time the (group x chunk) grid of k_iir_stages, iir_pieces_* the balanced-pieces kernel k_iir_pieces:
 * a non-zero launch count of a class proves which decomposition ran. */
int sgs_profile_enable(int on) {
    sgs::prof_collect();
    sgs::g_prof_on = on != 0;
    if (on) for (int i = 0; i < sgs::kProfCount; ++i) { sgs::g_prof_ms[i] = 0; sgs::g_prof_n[i] = 0; }
    return SGS_OK;
}
int sgs_profile_read(const char* name, double* total_ms, unsigned long long* launches) {
    SGS_ARG(name && total_ms && launches, "NULL argument");
    sgs::prof_collect();
    for (int i = 0; i < sgs::kProfCount; ++i)
        if (strcmp(name, sgs::kProfNames[i]) == 0) { *total_ms = sgs::g_prof_ms[i]; *launches = sgs::g_prof_n[i]; return SGS_OK; }
    sgs::set_error("unknown kernel class '%s'", name);
    return SGS_ERR_ARG;
}

int sgs_synchronize(void* stream) {
    SGS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return SGS_OK;
}

}  // extern "C"
