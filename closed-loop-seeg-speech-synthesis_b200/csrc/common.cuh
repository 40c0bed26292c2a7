// Shared helpers for libsgs (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SGS_OK 0
#define SGS_ERR_ARG 1
#define SGS_ERR_CUDA 2
#define SGS_ERR_UNSUPPORTED 3

namespace sgs {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SGS_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return sgs::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define SGS_ARG(cond, ...)                                                    \
    do {                                                                      \
        if (!(cond)) { sgs::set_error(__VA_ARGS__); return SGS_ERR_ARG; }     \
    } while (0)

// Launch accounting: bench.py reports how many of OUR kernels ran inside the timed region.
extern unsigned long long g_launches;
#define SGS_LAUNCHED() (++sgs::g_launches)

// A caller buffer may live on the host or on the device (cudaPointerGetAttributes decides).
// Host buffers are staged through device scratch on the call's stream.
bool is_device_ptr(const void* p);

struct Staged {
    void* dev = nullptr;      // device view of the buffer
    void* host = nullptr;     // original host pointer (nullptr when the caller passed device memory)
    size_t bytes = 0;
    bool owned = false;
};

int stage_in(Staged& s, const void* p, size_t bytes, cudaStream_t st);     // H2D if needed
int stage_out(Staged& s, void* p, size_t bytes, cudaStream_t st);          // allocate device mirror if needed
int finish_out(Staged& s, cudaStream_t st);                                // D2H if needed (async)
void release(Staged& s, cudaStream_t st);

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline leg).
enum ProfId { kProfIirInit = 0, kProfIirState, kProfIirCarry, kProfIirFeat, kProfStack, kProfLda, kProfGlBlocks, kProfGlOla,
              kProfLowpass, kProfStream, kProfGlBatch, kProfLogMel, kProfTrain, kProfLdaTc, kProfTrainTc, kProfPiecesState, kProfPiecesFeat, kProfLdaPack, kProfPiecesTail, kProfCount };
extern bool g_prof_on;
void prof_begin(int id, cudaStream_t st);
void prof_end(int id, cudaStream_t st);
struct ProfScope {
    int id; cudaStream_t st;
    ProfScope(int i, cudaStream_t s) : id(i), st(s) { if (g_prof_on) prof_begin(id, st); }
    ~ProfScope() { if (g_prof_on) prof_end(id, st); }
};

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so it
// is set once per DEVICE - a process that calls sgs_init() for a second device must not inherit the first one's flag - and
// its status is returned (a launch without the opt-in fails with invalid-value).  `done` = a static bit mask owned by the caller.
template <typename F>
static inline cudaError_t smem_optin(F* func, size_t bytes, unsigned long long* done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((*done >> dev) & 1ULL)) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && dev < 64) *done |= 1ULL << dev;
    return e;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace sgs
