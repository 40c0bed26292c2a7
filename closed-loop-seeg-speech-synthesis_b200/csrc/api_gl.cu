// C ABI for Griffin-Lim synthesis (include/sgs.h).
#include <math.h>
#include <vector>
#include "common.cuh"
#include "fft.cuh"
#include "../../include/sgs.h"

namespace sgs {
constexpr int kFft = 256, kHalf = 128, kHop = 160, kBlk = 480, kBins = 129;
constexpr int kLpMaxOrd = 8;
struct GlNodeTables { const double* window; const cplx* tw_half; const cplx* tw_full; const int* inv_idx; const double* inv_w; };
struct LpCoefs { double b[kLpMaxOrd + 1], a[kLpMaxOrd + 1]; int ord; };
int gl_blocks_run(const double* logmel, const double* noise, unsigned long long seed, double* blocks, const GlNodeTables& tab,
                  int n_sessions, int n_frames, int n_mels, int first_frame, int iters, cudaStream_t st);
int gl_emit_run(const double* blocks, const int* pos, const double* ola_window, double* v, double* states, double* zi,
                const double* phi, const LpCoefs& c, double norm_div, short* pcm, double* filtered, int n_sessions, int n_frames,
                int first_frame, long long n_out, int chunk, int n_chunks, cudaStream_t st);
}  // namespace sgs

struct sgs_gl_node {
    int n_mels = 0, iterations = 0, first_frame = 1;
    double norm_div = 1.01;
    sgs::LpCoefs lp;
    double *d_window = nullptr, *d_ola = nullptr, *d_inv_w = nullptr, *d_phi = nullptr;
    sgs::cplx *d_tw_half = nullptr, *d_tw_full = nullptr;
    int* d_inv_idx = nullptr;
    int lp_chunk = 0;
};

static cudaError_t upload(void** dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    return e;
}

extern "C" {

void sgs_gl_node_destroy(sgs_gl_node* n) {
    if (!n) return;
    cudaFree(n->d_window); cudaFree(n->d_ola); cudaFree(n->d_inv_w); cudaFree(n->d_phi);
    cudaFree(n->d_tw_half); cudaFree(n->d_tw_full); cudaFree(n->d_inv_idx);
    delete n;
}

int sgs_gl_node_create(sgs_gl_node** node, int fft_size, int hop, int block_len, int context_width, int n_mels,
                       const double* window, const double* ola_window, const int32_t* inv_idx, const double* inv_w,
                       const double* lp_b, const double* lp_a, int lp_order, const double* lp_phi, int lp_chunk,
                       double norm_div, int iterations) {
    using namespace sgs;
    SGS_ARG(node && window && ola_window && inv_idx && inv_w && lp_b && lp_a && lp_phi, "NULL argument");
    if (fft_size != kFft || hop != kHop || block_len != 3 || context_width != 1) {
        set_error("the node-semantics Griffin-Lim kernel is built for 16 ms frames / 10 ms shift at 16 kHz "
                  "(fft 256, hop 160, block 3, context 1); got fft %d hop %d block %d context %d", fft_size, hop, block_len, context_width);
        return SGS_ERR_UNSUPPORTED;
    }
    SGS_ARG(n_mels >= 1 && iterations >= 0 && lp_order >= 1 && lp_order <= kLpMaxOrd && lp_chunk >= 1, "bad configuration");
    SGS_ARG(lp_a[0] == 1.0, "low-pass denominator must be normalised (a[0] == 1)");
    sgs_gl_node* n = new sgs_gl_node();
    n->n_mels = n_mels; n->iterations = iterations; n->norm_div = norm_div; n->lp_chunk = lp_chunk;
    n->first_frame = block_len - context_width - 1;
    memset(&n->lp, 0, sizeof(n->lp));
    n->lp.ord = lp_order;
    for (int i = 0; i <= lp_order; ++i) { n->lp.b[i] = lp_b[i]; n->lp.a[i] = lp_a[i]; }
    for (int i = 0; i < kBins * 2; ++i)
        if (inv_idx[i] < 0 || inv_idx[i] >= n_mels) { delete n; set_error("inverse-mel tap index out of range"); return SGS_ERR_ARG; }
    std::vector<cplx> th(kHalf), tf(kBins);
    const double pi = 3.14159265358979323846;
    for (int t = 0; t < kHalf; ++t) th[t] = cplx{cos(2.0 * pi * t / kHalf), -sin(2.0 * pi * t / kHalf)};
    for (int k = 0; k < kBins; ++k) tf[k] = cplx{cos(2.0 * pi * k / kFft), -sin(2.0 * pi * k / kFft)};
    // exact values at the quadrant points keep the DC / Nyquist algebra free of 1e-17 leakage
    th[0] = cplx{1, 0}; th[kHalf / 4] = cplx{0, -1}; th[kHalf / 2] = cplx{-1, 0}; th[3 * kHalf / 4] = cplx{0, 1};
    tf[0] = cplx{1, 0}; tf[kFft / 4] = cplx{0, -1}; tf[kFft / 2] = cplx{-1, 0};
    cudaError_t e = upload((void**)&n->d_window, window, sizeof(double) * kFft);
    if (e == cudaSuccess) e = upload((void**)&n->d_ola, ola_window, sizeof(double) * kBlk);
    if (e == cudaSuccess) e = upload((void**)&n->d_inv_idx, inv_idx, sizeof(int) * kBins * 2);
    if (e == cudaSuccess) e = upload((void**)&n->d_inv_w, inv_w, sizeof(double) * kBins * 2);
    if (e == cudaSuccess) e = upload((void**)&n->d_tw_half, th.data(), sizeof(cplx) * kHalf);
    if (e == cudaSuccess) e = upload((void**)&n->d_tw_full, tf.data(), sizeof(cplx) * kBins);
    if (e == cudaSuccess) e = upload((void**)&n->d_phi, lp_phi, sizeof(double) * lp_order * lp_order);
    if (e != cudaSuccess) { sgs_gl_node_destroy(n); return cuda_fail(e, "table upload", __FILE__, __LINE__); }
    *node = n;
    return SGS_OK;
}

int sgs_gl_node_synthesize(sgs_gl_node* n, const double* logmel, int n_sessions, int n_frames, const int32_t* positions,
                           const double* noise, uint64_t seed, double* lp_state, int16_t* pcm, double* filtered,
                           double* blocks_out, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(n && n_sessions >= 1 && n_frames >= 0, "bad arguments");
    const int first = n->first_frame;
    if (n_frames <= first) return SGS_OK;
    SGS_ARG(logmel && positions && pcm, "NULL argument");
    for (int k = 1; k < n_frames; ++k)
        SGS_ARG(positions[k] > positions[k - 1] && positions[k] - positions[k - 1] <= 192, "bad write-head positions at frame %d", k);
    const long long n_out = (long long)positions[n_frames - 1] - (first > 0 ? positions[first - 1] : 0);
    const int chunk = n->lp_chunk;
    const int n_chunks = (int)((n_out + chunk - 1) / chunk);
    const int ord = n->lp.ord;

    Staged s_mel, s_noise, s_pcm, s_flt, s_blk;
    int* d_pos = nullptr;
    double *d_v = nullptr, *d_states = nullptr, *d_zi = nullptr;
    std::vector<double> zi_host((size_t)n_sessions * ord, 0.0);
    if (lp_state) memcpy(zi_host.data(), lp_state, sizeof(double) * n_sessions * ord);
    int rc = stage_in(s_mel, logmel, sizeof(double) * (size_t)n_sessions * n_frames * n->n_mels, st);
    if (rc == SGS_OK && noise) rc = stage_in(s_noise, noise, sizeof(double) * (size_t)n_sessions * n_frames * kBlk, st);
    if (rc == SGS_OK) rc = stage_out(s_pcm, pcm, sizeof(int16_t) * (size_t)n_sessions * n_out, st);
    if (rc == SGS_OK && filtered) rc = stage_out(s_flt, filtered, sizeof(double) * (size_t)n_sessions * n_out, st);
    if (rc == SGS_OK) rc = stage_out(s_blk, blocks_out, blocks_out ? sizeof(double) * (size_t)n_sessions * n_frames * kBlk : 0, st);
    double* d_blocks = (double*)s_blk.dev;
    bool own_blocks = false;
    cudaError_t e = cudaSuccess;
    if (rc == SGS_OK) {
        if (!d_blocks) { e = cudaMallocAsync((void**)&d_blocks, sizeof(double) * (size_t)n_sessions * n_frames * kBlk, st); own_blocks = true; }
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_pos, sizeof(int) * n_frames, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_v, sizeof(double) * (size_t)n_sessions * n_out, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_states, sizeof(double) * (size_t)n_sessions * n_chunks * kLpMaxOrd, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_zi, sizeof(double) * (size_t)n_sessions * ord, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_pos, positions, sizeof(int) * n_frames, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_zi, zi_host.data(), sizeof(double) * n_sessions * ord, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "scratch", __FILE__, __LINE__);
    }
    if (rc == SGS_OK) {
        GlNodeTables tab{n->d_window, n->d_tw_half, n->d_tw_full, n->d_inv_idx, n->d_inv_w};
        rc = gl_blocks_run((const double*)s_mel.dev, (const double*)s_noise.dev, seed, d_blocks, tab, n_sessions, n_frames,
                           n->n_mels, first, n->iterations, st);
    }
    if (rc == SGS_OK)
        rc = gl_emit_run(d_blocks, d_pos, n->d_ola, d_v, d_states, d_zi, n->d_phi, n->lp, n->norm_div, (short*)s_pcm.dev,
                         (double*)s_flt.dev, n_sessions, n_frames, first, n_out, chunk, n_chunks, st);
    if (rc == SGS_OK) rc = finish_out(s_pcm, st);
    if (rc == SGS_OK && filtered) rc = finish_out(s_flt, st);
    if (rc == SGS_OK && blocks_out) rc = finish_out(s_blk, st);
    if (rc == SGS_OK && lp_state) {
        e = cudaMemcpyAsync(zi_host.data(), d_zi, sizeof(double) * n_sessions * ord, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "state readback", __FILE__, __LINE__);
        else memcpy(lp_state, zi_host.data(), sizeof(double) * n_sessions * ord);
    }
    const bool sync = s_pcm.host || s_flt.host || s_blk.host;
    if (own_blocks && d_blocks) cudaFreeAsync(d_blocks, st);
    if (d_pos) cudaFreeAsync(d_pos, st);
    if (d_v) cudaFreeAsync(d_v, st);
    if (d_states) cudaFreeAsync(d_states, st);
    if (d_zi) cudaFreeAsync(d_zi, st);
    release(s_mel, st); release(s_noise, st); release(s_pcm, st); release(s_flt, st); release(s_blk, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"
