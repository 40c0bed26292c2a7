// C ABI for Griffin-Lim synthesis (include/sgs.h).
#include <math.h>
#include <algorithm>
#include <vector>
#include "kernels.cuh"
#include "../../include/sgs.h"


static void make_twiddles(int n, std::vector<sgs::cplx>& half, std::vector<sgs::cplx>& full) {
    // half[t] = exp(-2 pi i t / (n/2)), t < n/2 ; full[k] = exp(-2 pi i k / n), k <= n/2 ; exact at the quadrant points
    const double pi = 3.14159265358979323846;
    const int m = n / 2;
    half.resize(m); full.resize(m + 1);
    for (int t = 0; t < m; ++t) half[t] = sgs::cplx{cos(2.0 * pi * t / m), -sin(2.0 * pi * t / m)};
    for (int k = 0; k <= m; ++k) full[k] = sgs::cplx{cos(2.0 * pi * k / n), -sin(2.0 * pi * k / n)};
    half[0] = sgs::cplx{1, 0};
    if (m % 4 == 0) { half[m / 4] = sgs::cplx{0, -1}; half[m / 2] = sgs::cplx{-1, 0}; half[3 * m / 4] = sgs::cplx{0, 1}; }
    full[0] = sgs::cplx{1, 0}; full[m] = sgs::cplx{-1, 0};
    if (n % 4 == 0) full[n / 4] = sgs::cplx{0, -1};
}

struct sgs_gl_batch {
    int n_mels = 0;
    double *d_window = nullptr, *d_inv_w = nullptr;
    sgs::cplx *d_tw_half = nullptr, *d_tw_full = nullptr;
    int* d_inv_idx = nullptr;
};

struct sgs_gl_node {
    int n_mels = 0, iterations = 0, first_frame = 1, log_mels = 1;
    double norm_div = 1.01;
    sgs::LpCoefs lp;
    double *d_window = nullptr, *d_ola = nullptr, *d_inv_w = nullptr, *d_phi = nullptr, *d_phi_sub = nullptr;
    sgs::cplx *d_tw_full = nullptr, *d_tw_t = nullptr;
    int* d_inv_idx = nullptr;
    int lp_chunk = 0;
    int lp_carry_depth = 0;        // chunks after which Phi^depth is below 2^-70 (0: Phi is not small, sequential carry)
    // streaming state (sgs_gl_node_push): previous spectral frame + new ones, block ring, low-pass state
    double *d_mel = nullptr, *d_ring = nullptr, *d_lp = nullptr, *d_noise = nullptr;
    short* d_pcm = nullptr;
    long long frames_seen = 0;
    int ring_pos[sgs::kBlockRing];
    long long ring_index[sgs::kBlockRing];
    // cached write-head table of sgs_gl_node_synthesize (re-uploaded only when the caller's table changes: a copy of this
    // size goes through the H2D copy engine and would queue behind a concurrent bulk upload on another stream)
    std::vector<int32_t> h_pos;
    int* d_pos = nullptr;
    size_t d_pos_cap = 0;
    // scratch of sgs_gl_node_synthesize (blocks, overlap-added waveform, low-pass chunk states), one arena per stream, kept
    // between calls: at 32 sessions x 60 000 frames this is 10 GB, and taking it from the stream-ordered pool on every call
    // cost 6 ms of host time - 20 to 400 ms on the first call after a device synchronize, when the queue behind it is empty
    struct Arena { cudaStream_t st; char* p; size_t cap; };
    std::vector<Arena> arenas;
};

// the stream's arena, grown (never shrunk) to `bytes`; growing waits for the stream's earlier work
static cudaError_t gl_arena(sgs_gl_node* n, cudaStream_t st, size_t bytes, char** out) {
    sgs_gl_node::Arena* a = nullptr;
    for (auto& it : n->arenas) if (it.st == st) a = &it;
    if (!a) { n->arenas.push_back({st, nullptr, 0}); a = &n->arenas.back(); }
    if (a->cap < bytes) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        if (a->p) cudaFree(a->p);
        a->p = nullptr; a->cap = 0;
        e = cudaMalloc((void**)&a->p, bytes);
        if (e != cudaSuccess) return e;
        a->cap = bytes;
    }
    *out = a->p;
    return cudaSuccess;
}

static cudaError_t upload(void** dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    return e;
}

extern "C" {

void sgs_gl_node_destroy(sgs_gl_node* n) {
    if (!n) return;
    cudaFree(n->d_window); cudaFree(n->d_ola); cudaFree(n->d_inv_w); cudaFree(n->d_phi); cudaFree(n->d_phi_sub);
    cudaFree(n->d_tw_full); cudaFree(n->d_tw_t); cudaFree(n->d_inv_idx);
    cudaFree(n->d_mel); cudaFree(n->d_ring); cudaFree(n->d_lp); cudaFree(n->d_noise); cudaFree(n->d_pcm); cudaFree(n->d_pos);
    for (auto& a : n->arenas) cudaFree(a.p);
    delete n;
}

int sgs_gl_node_create(sgs_gl_node** node, int fft_size, int hop, int block_len, int context_width, int n_mels,
                       const double* window, const double* ola_window, const int32_t* inv_idx, const double* inv_w,
                       const double* lp_b, const double* lp_a, int lp_order, const double* lp_phi, int lp_chunk,
                       const double* lp_phi_sub, double norm_div, int iterations) {
    using namespace sgs;
    SGS_ARG(node && window && ola_window && inv_idx && inv_w && lp_b && lp_a && lp_phi && lp_phi_sub, "NULL argument");
    SGS_ARG(lp_chunk == 2048, "lp_chunk must be 2048 (32 sub-chunks of 64 samples)");
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
    {
        // |Phi|_inf^depth < 2^-70 bounds what the parallel carry drops (gl_node.cu:k_lp_carry_par)
        double norm = 0.0;
        for (int i = 0; i < lp_order; ++i) {
            double row = 0.0;
            for (int k = 0; k < lp_order; ++k) row += fabs(lp_phi[i * lp_order + k]);
            norm = std::max(norm, row);
        }
        n->lp_carry_depth = (norm > 0.0 && norm < 0.25) ? (int)ceil(-70.0 * log(2.0) / log(norm)) : (norm == 0.0 ? 1 : 0);
    }
    memset(&n->lp, 0, sizeof(n->lp));
    n->lp.ord = lp_order;
    for (int i = 0; i <= lp_order; ++i) { n->lp.b[i] = lp_b[i]; n->lp.a[i] = lp_a[i]; }
    for (int i = 0; i < kBins * 2; ++i)
        if (inv_idx[i] < 0 || inv_idx[i] >= n_mels) { delete n; set_error("inverse-mel tap index out of range"); return SGS_ERR_ARG; }
    std::vector<cplx> tf(kBins);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < kBins; ++k) tf[k] = cplx{cos(2.0 * pi * k / kFft), -sin(2.0 * pi * k / kFft)};
    // exact values at the quadrant points keep the DC / Nyquist algebra free of 1e-17 leakage
    tf[0] = cplx{1, 0}; tf[kFft / 4] = cplx{0, -1}; tf[kFft / 2] = cplx{-1, 0};
    for (int k = 1; k < kFft / 4; ++k) tf[kFft / 2 - k] = cplx{-tf[k].x, tf[k].y};      // w[128 - k] = -conj(w[k]) to the bit (gl_blocks8.cuh relies on it)
    // W128^(l k1) for the register-FFT kernel (gl_blocks8.cuh): [k1][l] with rows padded to 9 entries
    // second table: W128^((l + 48) k1) for the lanes of STFT frame 1, which hold point m in register slot (m + 10) mod 16
    std::vector<cplx> tt(2 * 16 * 9, cplx{0, 0});
    for (int f = 0; f < 2; ++f)
        for (int k1 = 0; k1 < 16; ++k1)
            for (int l = 0; l < 8; ++l) {
                const int ex = ((l + 48 * f) * k1) % kHalf;
                cplx w{cos(2.0 * pi * ex / kHalf), -sin(2.0 * pi * ex / kHalf)};
                if (ex == 0) w = cplx{1, 0};
                else if (ex == kHalf / 4) w = cplx{0, -1};
                else if (ex == kHalf / 2) w = cplx{-1, 0};
                else if (ex == 3 * kHalf / 4) w = cplx{0, 1};
                tt[(f * 16 + k1) * 9 + l] = w;
            }
    cudaError_t e = upload((void**)&n->d_window, window, sizeof(double) * kFft);
    if (e == cudaSuccess) e = upload((void**)&n->d_tw_t, tt.data(), sizeof(cplx) * tt.size());
    if (e == cudaSuccess) e = upload((void**)&n->d_ola, ola_window, sizeof(double) * kBlk);
    if (e == cudaSuccess) e = upload((void**)&n->d_inv_idx, inv_idx, sizeof(int) * kBins * 2);
    if (e == cudaSuccess) e = upload((void**)&n->d_inv_w, inv_w, sizeof(double) * kBins * 2);
    if (e == cudaSuccess) e = upload((void**)&n->d_tw_full, tf.data(), sizeof(cplx) * kBins);
    if (e == cudaSuccess) e = upload((void**)&n->d_phi, lp_phi, sizeof(double) * lp_order * lp_order);
    if (e == cudaSuccess) e = upload((void**)&n->d_phi_sub, lp_phi_sub, sizeof(double) * lp_order * lp_order);
    for (int i = 0; i < kBlockRing; ++i) { n->ring_pos[i] = 0; n->ring_index[i] = -1; }
    if (e == cudaSuccess) e = cudaMalloc((void**)&n->d_mel, sizeof(double) * (kMaxFramesPerPush + 1) * n_mels);
    if (e == cudaSuccess) e = cudaMalloc((void**)&n->d_ring, sizeof(double) * kBlockRing * kBlk);
    if (e == cudaSuccess) e = cudaMalloc((void**)&n->d_noise, sizeof(double) * (kMaxFramesPerPush + 1) * kBlk);
    if (e == cudaSuccess) e = cudaMalloc((void**)&n->d_pcm, sizeof(short) * kMaxFramesPerPush * 192);
    if (e == cudaSuccess) e = cudaMalloc((void**)&n->d_lp, sizeof(double) * kLpMaxOrd);
    if (e == cudaSuccess) e = cudaMemset(n->d_lp, 0, sizeof(double) * kLpMaxOrd);
    if (e == cudaSuccess) e = cudaMemset(n->d_mel, 0, sizeof(double) * (kMaxFramesPerPush + 1) * n_mels);
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
    if (n->h_pos.size() != (size_t)n_frames || memcmp(n->h_pos.data(), positions, sizeof(int32_t) * n_frames) != 0) {
        if (n->d_pos_cap < (size_t)n_frames) {
            SGS_CUDA(cudaStreamSynchronize(st));
            if (n->d_pos) SGS_CUDA(cudaFree(n->d_pos));
            n->d_pos = nullptr; n->d_pos_cap = 0; n->h_pos.clear();
            SGS_CUDA(cudaMalloc((void**)&n->d_pos, sizeof(int32_t) * n_frames));
            n->d_pos_cap = n_frames;
        }
        n->h_pos.assign(positions, positions + n_frames);
        SGS_CUDA(cudaMemcpyAsync(n->d_pos, n->h_pos.data(), sizeof(int32_t) * n_frames, cudaMemcpyHostToDevice, st));
    }
    int* d_pos = n->d_pos;
    double *d_v = nullptr, *d_states = nullptr, *d_zi = nullptr;
    std::vector<double> zi_host((size_t)n_sessions * ord, 0.0);
    if (lp_state) memcpy(zi_host.data(), lp_state, sizeof(double) * n_sessions * ord);
    int rc = stage_in(s_mel, logmel, sizeof(double) * (size_t)n_sessions * n_frames * n->n_mels, st);
    if (rc == SGS_OK && noise) rc = stage_in(s_noise, noise, sizeof(double) * (size_t)n_sessions * n_frames * kBlk, st);
    if (rc == SGS_OK) rc = stage_out(s_pcm, pcm, sizeof(int16_t) * (size_t)n_sessions * n_out, st);
    if (rc == SGS_OK && filtered) rc = stage_out(s_flt, filtered, sizeof(double) * (size_t)n_sessions * n_out, st);
    if (rc == SGS_OK) rc = stage_out(s_blk, blocks_out, blocks_out ? sizeof(double) * (size_t)n_sessions * n_frames * kBlk : 0, st);
    double* d_blocks = (double*)s_blk.dev;
    cudaError_t e = cudaSuccess;
    if (rc == SGS_OK) {
        auto pad = [](size_t b) { return (b + 255) / 256 * 256; };
        const size_t b_blocks = d_blocks ? 0 : pad(sizeof(double) * (size_t)n_sessions * n_frames * kBlk);
        const size_t b_v = pad(sizeof(double) * (size_t)n_sessions * n_out);
        const size_t b_states = pad(sizeof(double) * 2 * (size_t)n_sessions * n_chunks * kLpMaxOrd);
        const size_t b_zi = pad(sizeof(double) * 2 * (size_t)n_sessions * ord);                                        // in, out
        char* arena = nullptr;
        e = gl_arena(n, st, b_blocks + b_v + b_states + b_zi, &arena);
        if (e == cudaSuccess) {
            if (!d_blocks) d_blocks = (double*)arena;
            d_v = (double*)(arena + b_blocks);
            d_states = (double*)(arena + b_blocks + b_v);
            d_zi = (double*)(arena + b_blocks + b_v + b_states);
        }
        if (e == cudaSuccess)
            e = lp_state ? cudaMemcpyAsync(d_zi, zi_host.data(), sizeof(double) * n_sessions * ord, cudaMemcpyHostToDevice, st)
                         : cudaMemsetAsync(d_zi, 0, sizeof(double) * n_sessions * ord, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "scratch", __FILE__, __LINE__);
    }
    if (rc == SGS_OK) {
        GlNodeTables tab{n->d_window, n->d_tw_full, n->d_tw_t, n->d_inv_idx, n->d_inv_w, n->log_mels};
        rc = gl_blocks_run((const double*)s_mel.dev, (const double*)s_noise.dev, seed, d_blocks, tab, n_sessions, n_frames,
                           n->n_mels, first, n->iterations, 0, 0, st);
    }
    if (rc == SGS_OK)
        rc = gl_emit_run(d_blocks, d_pos, n->d_ola, d_v, d_states, d_zi, d_zi + (size_t)n_sessions * ord, n->lp_carry_depth, n->d_phi, n->lp, n->norm_div, (short*)s_pcm.dev,
                         (double*)s_flt.dev, n_sessions, n_frames, first, n_out, chunk, n_chunks, st);
    if (rc == SGS_OK) rc = finish_out(s_pcm, st);
    if (rc == SGS_OK && filtered) rc = finish_out(s_flt, st);
    if (rc == SGS_OK && blocks_out) rc = finish_out(s_blk, st);
    if (rc == SGS_OK && lp_state) {
        e = cudaMemcpyAsync(zi_host.data(), d_zi + (size_t)n_sessions * ord, sizeof(double) * n_sessions * ord, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "state readback", __FILE__, __LINE__);
        else memcpy(lp_state, zi_host.data(), sizeof(double) * n_sessions * ord);
    }
    const bool sync = s_pcm.host || s_flt.host || s_blk.host;
    release(s_mel, st); release(s_noise, st); release(s_pcm, st); release(s_flt, st); release(s_blk, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

void sgs_gl_batch_destroy(sgs_gl_batch* p) {
    if (!p) return;
    cudaFree(p->d_window); cudaFree(p->d_inv_w); cudaFree(p->d_tw_half); cudaFree(p->d_tw_full); cudaFree(p->d_inv_idx);
    delete p;
}

int sgs_gl_batch_create(sgs_gl_batch** plan, int win_len, int hop, int n_mels, const double* window, const int32_t* inv_idx,
                        const double* inv_w) {
    using namespace sgs;
    SGS_ARG(plan && window && inv_idx && inv_w, "NULL argument");
    if (win_len != 800 || hop != 160) {
        set_error("the batch Griffin-Lim kernel is built for 50 ms windows / 10 ms hop at 16 kHz (800 / 160); got %d / %d", win_len, hop);
        return SGS_ERR_UNSUPPORTED;
    }
    SGS_ARG(n_mels >= 1 && n_mels <= 64, "n_mels must be 1..64");
    const int bins = win_len / 2 + 1;
    for (int i = 0; i < bins * 2; ++i) SGS_ARG(inv_idx[i] >= 0 && inv_idx[i] < n_mels, "inverse-mel tap index out of range");
    sgs_gl_batch* p = new sgs_gl_batch();
    p->n_mels = n_mels;
    std::vector<cplx> th, tf;
    make_twiddles(win_len, th, tf);
    cudaError_t e = upload((void**)&p->d_window, window, sizeof(double) * win_len);
    if (e == cudaSuccess) e = upload((void**)&p->d_inv_idx, inv_idx, sizeof(int) * bins * 2);
    if (e == cudaSuccess) e = upload((void**)&p->d_inv_w, inv_w, sizeof(double) * bins * 2);
    if (e == cudaSuccess) e = upload((void**)&p->d_tw_half, th.data(), sizeof(cplx) * th.size());
    if (e == cudaSuccess) e = upload((void**)&p->d_tw_full, tf.data(), sizeof(cplx) * tf.size());
    if (e != cudaSuccess) { sgs_gl_batch_destroy(p); return cuda_fail(e, "table upload", __FILE__, __LINE__); }
    *plan = p;
    return SGS_OK;
}

int sgs_gl_batch_synthesize(sgs_gl_batch* p, const double* logmel, int n_utt, int n_frames, const double* noise,
                            int64_t noise_stride, int iterations, int16_t* pcm, double* waveform, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(p && logmel && noise && pcm, "NULL argument");
    SGS_ARG(n_utt >= 1 && iterations >= 0, "bad arguments");
    SGS_ARG(n_frames >= 6, "need at least 6 spectral frames (the reference divides by max|x| = 0 below that)");
    const long long x_len = 160LL * (n_frames - 1) + 800, n_out = 160LL * n_frames;
    SGS_ARG(noise_stride >= x_len, "noise rows must hold at least %lld samples", x_len);
    Staged s_mel, s_noise, s_pcm, s_wave;
    double *d_x = nullptr, *d_mx = nullptr;
    int rc = stage_in(s_mel, logmel, sizeof(double) * (size_t)n_utt * n_frames * p->n_mels, st);
    if (rc == SGS_OK) rc = stage_in(s_noise, noise, sizeof(double) * ((size_t)(n_utt - 1) * noise_stride + x_len), st);
    if (rc == SGS_OK) rc = stage_out(s_pcm, pcm, sizeof(int16_t) * (size_t)n_utt * n_out, st);
    if (rc == SGS_OK && waveform) rc = stage_out(s_wave, waveform, sizeof(double) * (size_t)n_utt * n_out, st);
    cudaError_t e = cudaSuccess;
    if (rc == SGS_OK) {
        e = cudaMallocAsync((void**)&d_x, sizeof(double) * (size_t)n_utt * x_len, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_mx, sizeof(double) * n_utt, st);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(d_x, sizeof(double) * x_len, s_noise.dev, sizeof(double) * noise_stride, sizeof(double) * x_len,
                                  n_utt, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "scratch", __FILE__, __LINE__);
    }
    if (rc == SGS_OK) {
        GlBatchTables tab{p->d_window, p->d_tw_half, p->d_tw_full, p->d_inv_idx, p->d_inv_w};
        rc = gl_batch_run((const double*)s_mel.dev, d_x, tab, n_utt, n_frames, p->n_mels, iterations, x_len, d_mx, (short*)s_pcm.dev, st);
    }
    if (rc == SGS_OK && waveform) {
        e = cudaMemcpy2DAsync(s_wave.dev, sizeof(double) * n_out, d_x, sizeof(double) * x_len, sizeof(double) * n_out, n_utt,
                              cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "waveform copy", __FILE__, __LINE__);
    }
    if (rc == SGS_OK) rc = finish_out(s_pcm, st);
    if (rc == SGS_OK && waveform) rc = finish_out(s_wave, st);
    const bool sync = s_pcm.host || s_wave.host;
    if (d_x) cudaFreeAsync(d_x, st);
    if (d_mx) cudaFreeAsync(d_mx, st);
    release(s_mel, st); release(s_noise, st); release(s_pcm, st); release(s_wave, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

int sgs_logmel(const double* audio, int64_t n_audio, const double* window, int win_len, int shift, const double* mel,
               int n_bins, int n_mels, int64_t n_frames, double* out, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(audio && window && mel && out, "NULL argument");
    if (win_len != 256 || n_bins != 129) {
        set_error("the log-mel kernel is built for 16 ms windows at 16 kHz (256 samples, 129 bins); got %d / %d", win_len, n_bins);
        return SGS_ERR_UNSUPPORTED;
    }
    SGS_ARG(shift >= 1 && shift <= win_len && n_mels >= 1 && n_frames >= 0, "bad arguments");
    if (n_frames == 0) return SGS_OK;
    std::vector<cplx> th, tf;
    make_twiddles(win_len, th, tf);
    Staged s_a, s_w, s_m, s_th, s_tf, s_o;
    int rc = stage_in(s_a, audio, sizeof(double) * (size_t)n_audio, st);
    if (rc == SGS_OK) rc = stage_in(s_w, window, sizeof(double) * win_len, st);
    if (rc == SGS_OK) rc = stage_in(s_m, mel, sizeof(double) * (size_t)n_bins * n_mels, st);
    if (rc == SGS_OK) rc = stage_in(s_th, th.data(), sizeof(cplx) * th.size(), st);
    if (rc == SGS_OK) rc = stage_in(s_tf, tf.data(), sizeof(cplx) * tf.size(), st);
    if (rc == SGS_OK) rc = stage_out(s_o, out, sizeof(double) * (size_t)n_frames * n_mels, st);
    if (rc == SGS_OK)
        rc = logmel_run((const double*)s_a.dev, n_audio, (const double*)s_w.dev, (const cplx*)s_th.dev, (const cplx*)s_tf.dev,
                        (const double*)s_m.dev, n_mels, n_frames, shift, win_len - shift, (double*)s_o.dev, st);
    if (rc == SGS_OK) rc = finish_out(s_o, st);
    // th/tf are pageable host vectors: their H2D copies were staged synchronously, safe to drop
    release(s_a, st); release(s_w, st); release(s_m, st); release(s_th, st); release(s_tf, st);
    const bool sync = true;
    release(s_o, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"

namespace sgs {
/* Everything of a node push except the read-back: logmel / noise may be host or device memory; the int16 hop(s) land in
 * d_pcm (or the node's own buffer when NULL); *n_pcm = number of samples emitted. */
int gl_node_enqueue(sgs_gl_node* s, const double* logmel, int n, const int32_t* pos, int32_t pos_before, const double* noise,
                    uint64_t seed, short* d_pcm, int* n_pcm, cudaStream_t st, bool src_pinned) {
    SGS_ARG(s && logmel && pos && n_pcm, "NULL argument");
    SGS_ARG(n >= 1 && n <= kMaxFramesPerPush, "push takes 1..%d frames (got %d)", kMaxFramesPerPush, n);
    const int nm = s->n_mels, first = s->first_frame;
    const long long k0 = s->frames_seen;
    // row 0 of d_mel holds the previous frame; append the new ones behind it
    SGS_CUDA(cudaMemcpyAsync(s->d_mel + nm, logmel, sizeof(double) * n * nm, cudaMemcpyDefault, st));
    if (noise) { const int rcn = copy_in_small(s->d_noise + kBlk, noise, sizeof(double) * n * kBlk, src_pinned, st); if (rcn != SGS_OK) return rcn; }
    EmitFrames fr;
    memset(&fr, 0, sizeof(fr));
    int total = 0, prev = pos_before;
    // frames k0 .. k0+n-1; a block exists for k >= first (the node returns early before that, GriffinLim.py:131)
    const int skip = (k0 < first) ? (int)std::min<long long>(first - k0, n) : 0;
    for (int i = 0; i < n; ++i) {
        SGS_ARG(pos[i] > prev && pos[i] - prev <= 192, "bad write-head position at new frame %d", i);
        if (i >= skip) {
            const long long k = k0 + i;
            const int q = fr.n++;
            fr.index[q] = k; fr.pos[q] = pos[i]; fr.prev[q] = prev;
            total += pos[i] - prev;
            s->ring_pos[k & (kBlockRing - 1)] = pos[i];
            s->ring_index[k & (kBlockRing - 1)] = k;
        }
        prev = pos[i];
    }
    for (int i = 0; i < kBlockRing; ++i) { fr.ring_pos[i] = s->ring_pos[i]; fr.ring_index[i] = s->ring_index[i]; }
    int rc = SGS_OK;
    if (fr.n > 0) {
        GlNodeTables tab{s->d_window, s->d_tw_full, s->d_tw_t, s->d_inv_idx, s->d_inv_w, s->log_mels};
        // local frame j (row j of d_mel) is running frame k0 - 1 + j; blocks for local frames [1 + skip, n]
        rc = gl_blocks_run(s->d_mel, noise ? s->d_noise : nullptr, seed, s->d_ring, tab, 1, n + 1, nm, 1 + skip, s->iterations,
                           k0 - 1, kBlockRing, st);
        if (rc == SGS_OK)
            rc = gl_emit_stream_run(s->d_ring, s->d_ola, s->d_lp, d_pcm ? d_pcm : s->d_pcm, s->lp, s->norm_div, first, fr, st);
    }
    // the newest frame becomes the "previous" one
    SGS_CUDA(cudaMemcpyAsync(s->d_mel, s->d_mel + (size_t)n * nm, sizeof(double) * nm, cudaMemcpyDeviceToDevice, st));
    s->frames_seen += n;
    *n_pcm = total;
    return rc;
}
}  // namespace sgs

extern "C" {

int sgs_gl_node_set_log_mels(sgs_gl_node* s, int log_mels) {
    SGS_ARG(s, "NULL argument");
    s->log_mels = log_mels ? 1 : 0;
    return SGS_OK;
}

int sgs_gl_node_rebase(sgs_gl_node* s, int32_t delta) {
    SGS_ARG(s, "NULL argument");
    for (int i = 0; i < sgs::kBlockRing; ++i) s->ring_pos[i] -= delta;
    return SGS_OK;
}

/* Streaming form: feed `n` new spectral frames (host), get the audio the node would have emitted for them. */
int sgs_gl_node_push(sgs_gl_node* s, const double* logmel, int n, const int32_t* pos, int32_t pos_before, const double* noise,
                     uint64_t seed, int16_t* pcm, int* n_pcm, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(pcm, "NULL argument");
    int rc = gl_node_enqueue(s, logmel, n, pos, pos_before, noise, seed, nullptr, n_pcm, st);
    if (rc == SGS_OK && *n_pcm > 0) SGS_CUDA(cudaMemcpyAsync(pcm, s->d_pcm, sizeof(short) * *n_pcm, cudaMemcpyDeviceToHost, st));
    if (rc == SGS_OK) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

/* Test hook: out[i] = exp(angle(re[i] + 1j*im[i])) as k_gl_blocks evaluates it (GriffinLim.py:93). */
int sgs_exp_angle(const double* im, const double* re, int64_t n, double* out, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(im && re && out && n >= 0, "bad arguments");
    Staged si, sr, so;
    int rc = stage_in(si, im, sizeof(double) * (size_t)n, st);
    if (rc == SGS_OK) rc = stage_in(sr, re, sizeof(double) * (size_t)n, st);
    if (rc == SGS_OK) rc = stage_out(so, out, sizeof(double) * (size_t)n, st);
    if (rc == SGS_OK) rc = exp_angle_run((const double*)si.dev, (const double*)sr.dev, n, (double*)so.dev, st);
    if (rc == SGS_OK) rc = finish_out(so, st);
    const bool sync = so.host != nullptr;
    release(si, st); release(sr, st); release(so, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

int sgs_dequantize(const double* medians, int n_bins, int n_levels, const double* taps, int radius, const double* labels,
                   int64_t n_rows, int smooth, double* out, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(medians && n_bins >= 1 && n_levels >= 1 && n_rows >= 0, "bad arguments");
    SGS_ARG(!smooth || (taps && radius >= 1 && radius < n_bins), "smoothing taps missing");
    if (n_rows == 0) return SGS_OK;
    SGS_ARG(labels && out, "NULL argument");
    Staged sm, stp, sl, so;
    int rc = stage_in(sm, medians, sizeof(double) * n_bins * n_levels, st);
    if (rc == SGS_OK && smooth) rc = stage_in(stp, taps, sizeof(double) * (2 * radius + 1), st);
    if (rc == SGS_OK) rc = stage_in(sl, labels, sizeof(double) * (size_t)n_rows * n_bins, st);
    if (rc == SGS_OK) rc = stage_out(so, out, sizeof(double) * (size_t)n_rows * n_bins, st);
    if (rc == SGS_OK)
        rc = dequantize_run((const double*)sl.dev, (const double*)sm.dev, (const double*)stp.dev, radius, smooth, n_bins, n_levels,
                            n_rows, (double*)so.dev, st);
    if (rc == SGS_OK) rc = finish_out(so, st);
    const bool sync = so.host != nullptr;
    release(sm, st); release(stp, st); release(sl, st); release(so, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"
