// C ABI for feature extraction (include/sgs.h).
#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <vector>
#include "kernels.cuh"
#include "../../include/sgs.h"


struct sgs_feat_plan {
    int n_biquads = 0;
    bool monic = false;
    int zero_fill = 0;
    sgs::FeatCoefs cf;
    double* d_zf = nullptr;               // zero-fill response table on the device
    double* d_coef = nullptr;             // [n_biquads][5] on the device (per-lane register fill)
    // cached window table (re-uploaded only when the caller's table changes)
    std::vector<int32_t> h_starts;
    int32_t* d_starts = nullptr;
    size_t d_starts_cap = 0;
    // modal tail of the warm-up (sgs_feat_plan_set_tail)
    bool has_tail = false;
    sgs::TailTab tail;
    double* d_tail_matrix = nullptr;      // [2 n_biquads][2 n_modes]
};

extern "C" {

int sgs_feat_plan_create(sgs_feat_plan** plan, int n_filters, const double* coef, const double* zi_unit,
                         const double* zi_last_warm, const double* zero_fill_response, int zero_fill) {
    SGS_ARG(plan && coef && zi_unit && zi_last_warm, "NULL argument");
    SGS_ARG(n_filters == 2 || n_filters == 3, "n_filters must be 2 or 3 (got %d)", n_filters);
    SGS_ARG(zero_fill >= 0 && (zero_fill == 0 || zero_fill_response), "zero_fill table missing");
    sgs_feat_plan* p = new sgs_feat_plan();
    p->n_biquads = n_filters * sgs::kSecPerFilter;
    p->zero_fill = zero_fill;
    memset(&p->cf, 0, sizeof(p->cf));
    bool monic = true;
    for (int i = 0; i < p->n_biquads; ++i) {
        for (int k = 0; k < 5; ++k) p->cf.c[i][k] = coef[i * 5 + k];
        p->cf.zi[i][0] = zi_unit[i * 2];
        p->cf.zi[i][1] = zi_unit[i * 2 + 1];
        // monic sections (b0 = b2 = 1) take the 4-flop form.  scipy's zpk2sos leaves b2 = 1 - 2^-52 on some
        // notch sections (2048 Hz); a coefficient within 4 ulp of 1 is treated as 1 (documented in DESIGN.md).
        const double ulp4 = 4 * 2.220446049250313e-16;
        if (i % sgs::kSecPerFilter != 0 && !(fabs(coef[i * 5] - 1.0) <= ulp4 && fabs(coef[i * 5 + 2] - 1.0) <= ulp4)) monic = false;
    }
    p->monic = monic;
    for (int s = 0; s < sgs::kSecPerFilter; ++s) {
        p->cf.zi_warm[s][0] = zi_last_warm[s * 2];
        p->cf.zi_warm[s][1] = zi_last_warm[s * 2 + 1];
    }
    if (zero_fill > 0) {
        cudaError_t e = cudaMalloc(&p->d_zf, sizeof(double) * zero_fill);
        if (e == cudaSuccess) e = cudaMemcpy(p->d_zf, zero_fill_response, sizeof(double) * zero_fill, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { delete p; return sgs::cuda_fail(e, "zero-fill table upload", __FILE__, __LINE__); }
    }
    {
        cudaError_t e = cudaMalloc(&p->d_coef, sizeof(double) * 5 * p->n_biquads);
        if (e == cudaSuccess) e = cudaMemcpy(p->d_coef, coef, sizeof(double) * 5 * p->n_biquads, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { sgs_feat_plan_destroy(p); return sgs::cuda_fail(e, "coefficient upload", __FILE__, __LINE__); }
    }
    *plan = p;
    return SGS_OK;
}

void sgs_feat_plan_destroy(sgs_feat_plan* p) {
    if (!p) return;
    if (p->d_zf) cudaFree(p->d_zf);
    if (p->d_coef) cudaFree(p->d_coef);
    if (p->d_starts) cudaFree(p->d_starts);
    if (p->d_tail_matrix) cudaFree(p->d_tail_matrix);
    delete p;
}

int sgs_feat_plan_set_tail(sgs_feat_plan* p, int near_len, int n_modes, const int32_t* mode_len, const double* lam,
                           const int32_t* warp_blocks, const double* warp_shift, const double* state_matrix, const double* kappa) {
    using namespace sgs;
    SGS_ARG(p, "NULL plan");
    if (n_modes == 0) { p->has_tail = false; return SGS_OK; }
    SGS_ARG(mode_len && lam && warp_blocks && warp_shift && state_matrix && kappa, "NULL argument");
    SGS_ARG(n_modes > 0 && n_modes <= kTailMaxModes && n_modes % 4 == 0, "n_modes must be 4, 8, 12 or 16 (got %d)", n_modes);
    SGS_ARG(near_len >= 64 && near_len % 64 == 0, "near_len must be a positive multiple of 64 (got %d)", near_len);
    TailTab& t = p->tail;
    memset(&t, 0, sizeof(t));
    for (int m = 0; m < n_modes; ++m) {
        SGS_ARG(mode_len[m] >= 0 && mode_len[m] % kTailBlock == 0, "mode_len[%d] = %d is not a multiple of %d", m, mode_len[m], kTailBlock);
        SGS_ARG(m == 0 || mode_len[m] <= mode_len[m - 1], "modes must be sorted by length, longest first");
        SGS_ARG(m % 4 == 0 || mode_len[m] == mode_len[m - 1], "modes join in groups of 4 of equal length");
        t.mode_blocks[m] = mode_len[m] / kTailBlock;
        t.lam[m][0] = lam[2 * m];
        t.lam[m][1] = lam[2 * m + 1];
        t.rec[m][0] = 2.0 * lam[2 * m];
        t.rec[m][1] = -(lam[2 * m] * lam[2 * m] + lam[2 * m + 1] * lam[2 * m + 1]);
        for (int w = 0; w < kTailWarps; ++w) {
            t.shift[w][m][0] = warp_shift[(w * n_modes + m) * 2];
            t.shift[w][m][1] = warp_shift[(w * n_modes + m) * 2 + 1];
        }
    }
    // every block of every group belongs to exactly one warp
    for (int gq = 0; gq < kTailMaxModes / 4; ++gq) {
        const int blocks = gq * 4 < n_modes ? t.mode_blocks[gq * 4] : 0;
        std::vector<char> seen(blocks, 0);
        for (int w = 0; w < kTailWarps; ++w) {
            const int lo = warp_blocks[(w * (kTailMaxModes / 4) + gq) * 2], hi = warp_blocks[(w * (kTailMaxModes / 4) + gq) * 2 + 1];
            SGS_ARG(lo >= 0 && lo <= hi && hi <= blocks, "warp %d, group %d: blocks [%d, %d) outside [0, %d)", w, gq, lo, hi, blocks);
            for (int d = lo; d < hi; ++d) { SGS_ARG(!seen[d], "block %d of group %d dealt twice", d, gq); seen[d] = 1; }
            t.blk_lo[w][gq] = lo;
            t.blk_hi[w][gq] = hi;
        }
        for (int d = 0; d < blocks; ++d) SGS_ARG(seen[d], "block %d of group %d dealt to no warp", d, gq);
    }
    t.n_modes = n_modes;
    t.near_len = near_len;
    t.far_len = mode_len[0];
    const size_t bytes = sizeof(double) * 2 * p->n_biquads * 2 * n_modes;                 // state matrix, then kappa: the same size
    if (p->d_tail_matrix) { cudaFree(p->d_tail_matrix); p->d_tail_matrix = nullptr; }
    cudaError_t e = cudaMalloc(&p->d_tail_matrix, 2 * bytes);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_tail_matrix, state_matrix, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((char*)p->d_tail_matrix + bytes, kappa, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { p->has_tail = false; return cuda_fail(e, "tail matrix upload", __FILE__, __LINE__); }
    p->has_tail = true;
    return SGS_OK;
}

int sgs_feat_extract(sgs_feat_plan* p, const void* x, int x_is_f64, int64_t n_samples, int n_channels, int n_sessions,
                     int64_t session_stride, const int32_t* win_starts, int n_windows, int window_len, int n_chunks,
                     int64_t chunk_len, int horizon, const double* phi, double* feat, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(p && x && feat, "NULL argument");
    SGS_ARG(n_samples >= 1 && n_channels >= 1 && n_sessions >= 1, "empty input");
    SGS_ARG(n_windows >= 0 && window_len >= 1, "bad window table");
    if (n_windows == 0) return SGS_OK;
    SGS_ARG(win_starts != nullptr, "win_starts is NULL");
    SGS_ARG(n_chunks >= 1 && (n_chunks == 1 || (chunk_len >= 1 && (int64_t)(n_chunks - 1) * chunk_len < n_samples)),
            "bad chunk plan (%d chunks of %lld over %lld samples)", n_chunks, (long long)chunk_len, (long long)n_samples);
    const bool apply_phi = n_chunks > 2 && horizon >= chunk_len;
    SGS_ARG(!apply_phi || phi != nullptr, "phi = A^chunk_len is required when horizon >= chunk_len");
    SGS_ARG(n_chunks == 1 || horizon >= 1, "horizon must be >= 1");
    if (session_stride == 0) session_stride = n_samples * n_channels;
    // validate the window table on the host (it is O(frames), tiny next to the sample data)
    for (int k = 0; k < n_windows; ++k) {
        SGS_ARG(k == 0 || win_starts[k] > win_starts[k - 1], "win_starts must be strictly increasing (k=%d)", k);
    }
    SGS_ARG(win_starts[0] >= -p->zero_fill, "first window starts before the zero-fill region");
    SGS_ARG((int64_t)win_starts[n_windows - 1] + window_len <= n_samples, "last window ends past the input");
    const int t_first = win_starts[0] < 0 ? win_starts[0] : 0;

    // chunk bounds + window ownership (by window start)
    std::vector<long long> bounds(n_chunks + 1);
    for (int j = 0; j < n_chunks; ++j) bounds[j] = (long long)j * chunk_len;
    bounds[n_chunks] = n_samples;
    std::vector<int> kfirst(n_chunks + 1);
    {
        int k = 0;
        kfirst[0] = 0;
        for (int j = 1; j < n_chunks; ++j) {
            while (k < n_windows && win_starts[k] < bounds[j]) ++k;
            kfirst[j] = k;
        }
        kfirst[n_chunks] = n_windows;
    }

    // window table upload (cached)
    if (p->h_starts.size() != (size_t)n_windows || memcmp(p->h_starts.data(), win_starts, sizeof(int32_t) * n_windows) != 0) {
        if (p->d_starts_cap < (size_t)n_windows) {
            SGS_CUDA(cudaStreamSynchronize(st));
            if (p->d_starts) SGS_CUDA(cudaFree(p->d_starts));
            SGS_CUDA(cudaMalloc(&p->d_starts, sizeof(int32_t) * n_windows));
            p->d_starts_cap = n_windows;
        }
        p->h_starts.assign(win_starts, win_starts + n_windows);
        SGS_CUDA(cudaMemcpyAsync(p->d_starts, p->h_starts.data(), sizeof(int32_t) * n_windows, cudaMemcpyHostToDevice, st));
    }

    FeatGeom g;
    g.n_samples = n_samples;
    g.n_channels = n_channels;
    g.n_sessions = n_sessions;
    g.session_stride = session_stride;
    g.n_streams = n_channels * n_sessions;
    g.n_windows = n_windows;
    g.window_len = window_len;
    g.t_first = t_first;
    g.zero_fill = p->zero_fill;
    g.n_chunks = n_chunks;
    g.horizon = horizon;
    g.state_stride = g.n_streams;

    const size_t esz = x_is_f64 ? 8 : 4;
    const size_t x_bytes = ((size_t)(n_sessions - 1) * session_stride + (size_t)n_samples * n_channels) * esz;
    const size_t f_bytes = (size_t)n_sessions * n_windows * n_channels * sizeof(double);
    Staged sx, sf;
    int rc = stage_in(sx, x, x_bytes, st);
    if (rc) return rc;
    rc = stage_out(sf, feat, f_bytes, st);
    if (rc) { release(sx, st); return rc; }

    const int ns = 2 * p->n_biquads;
    // ---- balanced pieces (feat.cu:k_iir_pieces) when the zero-state horizon is short against the work of one CTA ---------------------
    {
        const long long G = (g.n_streams + 31) / 32;
        int sms = 148;
        { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        const char* env_p = getenv("SGS_FEAT_PIECES_P");                     // tests force a small piece count on small inputs
        const long long P = (env_p && atoi(env_p) > 0) ? atoi(env_p) : 4LL * sms;   // one CTA per SM, one pipeline per scheduler (feat.cu)
        long long piece_len = (G * n_samples + P - 1) / P;
        piece_len = (piece_len + 63) / 64 * 64;
        const char* env = getenv("SGS_FEAT_PIECES");
        const bool want = !(env && env[0] == '0');
        // a cut costs a warm-up: pieces of at least twice the zero-state horizon, or - with the modal tail, whose cascade run is
        // near_len samples - of at least 4 near_len (SGS_FEAT_PIECES_MINLEN overrides, for measurements)
        long long min_len = p->has_tail ? 4LL * p->tail.near_len : 2LL * horizon;
        if (const char* env_m = getenv("SGS_FEAT_PIECES_MINLEN")) { if (atoll(env_m) > 0) min_len = atoll(env_m); }
        if (want && !apply_phi && n_chunks > 1 && horizon % 64 == 0 && piece_len >= min_len && piece_len <= n_samples) {
            std::vector<FeatSeg> segs;
            std::vector<int> piece_first;
            auto first_window_at = [&](long long t) {                        // first window whose start is >= t
                return (int)(std::lower_bound(win_starts, win_starts + n_windows, t, [](int32_t a, long long b) { return (long long)a < b; }) - win_starts);
            };
            const long long total = G * n_samples;
            for (long long b0 = 0; b0 < total; b0 += piece_len) {
                piece_first.push_back((int)segs.size());
                const long long b1 = std::min(total, b0 + piece_len);
                long long pos = b0;
                while (pos < b1) {
                    const long long grp = pos / n_samples;
                    long long t0 = pos - grp * n_samples;
                    t0 = t0 / 64 * 64;                                        // segment starts on a batch boundary of its recording
                    long long t1 = std::min(b1, (grp + 1) * n_samples) - grp * n_samples;
                    const bool to_end = (grp + 1) * n_samples <= b1;
                    if (!to_end) t1 = t1 / 64 * 64;
                    FeatSeg sg;
                    sg.group = (int)grp;
                    sg.t_begin = t0;
                    sg.warm_begin = t0 > horizon ? t0 - horizon : 0;
                    sg.tail = 0;
                    if (p->has_tail && t0 >= 2LL * p->tail.near_len) {
                        // far past as modal sums, cascade over near_len only; 2: the recording starts inside the tail's
                        // horizon, so the initial state still counts (k_iir_tail)
                        sg.tail = t0 >= (long long)p->tail.near_len + p->tail.far_len ? 1 : 2;
                        sg.warm_begin = t0 - p->tail.near_len;
                    }
                    sg.k_lo = t0 == 0 ? 0 : first_window_at(t0);
                    sg.k_hi = to_end ? n_windows : first_window_at(t1);
                    if (t1 > t0) segs.push_back(sg);
                    pos = grp * n_samples + (to_end ? n_samples : std::max(t1, t0 + 1));
                    if (!to_end) break;                                       // the rest of this group belongs to the next piece
                }
            }
            piece_first.push_back((int)segs.size());
            const int n_pieces = (int)piece_first.size() - 1;
            const size_t seg_bytes = sizeof(FeatSeg) * segs.size(), pf_bytes = sizeof(int) * piece_first.size();
            const size_t init_bytes = sizeof(double) * (size_t)ns * g.n_streams, slot_bytes = sizeof(double) * (size_t)ns * 32 * segs.size();
            char* d_tab2 = nullptr;
            double* d_state = nullptr;
            const size_t pf_off = (seg_bytes + 15) & ~(size_t)15;
            cudaError_t e2 = cudaMallocAsync((void**)&d_tab2, pf_off + pf_bytes, st);
            if (e2 == cudaSuccess) e2 = cudaMallocAsync((void**)&d_state, init_bytes + slot_bytes, st);
            if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(d_tab2, segs.data(), seg_bytes, cudaMemcpyHostToDevice, st);
            if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(d_tab2 + pf_off, piece_first.data(), pf_bytes, cudaMemcpyHostToDevice, st);
            if (e2 != cudaSuccess) { release(sx, st); release(sf, st); return cuda_fail(e2, "piece tables", __FILE__, __LINE__); }
            // (both tables are far below the 64 KB that the runtime stages synchronously, so the vectors may go out of scope)
            rc = feat_run_pieces(p->n_biquads, p->monic, sx.dev, x_is_f64 != 0, (double*)sf.dev, d_state, d_state + (size_t)ns * g.n_streams,
                                 (const FeatSeg*)d_tab2, (const int*)(d_tab2 + pf_off), n_pieces, (int)segs.size(),
                                 p->has_tail ? &p->tail : nullptr, p->d_tail_matrix, p->d_starts, p->d_zf, p->cf, g, st);
            if (rc == SGS_OK) rc = finish_out(sf, st);
            cudaFreeAsync(d_tab2, st);
            cudaFreeAsync(d_state, st);
            const bool sync2 = sf.host != nullptr;
            release(sx, st);
            release(sf, st);
            if (rc == SGS_OK && sync2) SGS_CUDA(cudaStreamSynchronize(st));
            return rc;
        }
    }
    // small per-call tables + carry scratch, stream-ordered
    const size_t tab_bytes = sizeof(long long) * (n_chunks + 1) + sizeof(int) * (n_chunks + 1) + (apply_phi ? sizeof(double) * ns * ns : 0);
    const size_t carry_bytes = sizeof(double) * (size_t)n_chunks * ns * g.n_streams;      // state slot per chunk start
    char* d_tab = nullptr;
    double* d_carry = nullptr;
    std::vector<char> h_tab(tab_bytes);
    size_t off_k = sizeof(long long) * (n_chunks + 1), off_phi = off_k + sizeof(int) * (n_chunks + 1);
    off_phi = (off_phi + 7) & ~(size_t)7;
    h_tab.resize(off_phi + (apply_phi ? sizeof(double) * ns * ns : 0));
    memcpy(h_tab.data(), bounds.data(), sizeof(long long) * (n_chunks + 1));
    memcpy(h_tab.data() + off_k, kfirst.data(), sizeof(int) * (n_chunks + 1));
    if (apply_phi) memcpy(h_tab.data() + off_phi, phi, sizeof(double) * ns * ns);
    cudaError_t e = cudaMallocAsync((void**)&d_tab, h_tab.size(), st);
    if (e == cudaSuccess && carry_bytes) e = cudaMallocAsync((void**)&d_carry, carry_bytes, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab, h_tab.data(), h_tab.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { release(sx, st); release(sf, st); return cuda_fail(e, "scratch", __FILE__, __LINE__); }
    // the pageable h_tab copy above is staged synchronously by the runtime, so h_tab may go out of scope

    rc = feat_run(p->n_biquads, p->monic, sx.dev, x_is_f64 != 0, (double*)sf.dev, d_carry, (const double*)(d_tab + off_phi),
                  apply_phi, (const long long*)d_tab, (const int*)(d_tab + off_k), p->d_starts, p->d_zf, p->d_coef, p->cf, g, st);
    if (rc == SGS_OK) rc = finish_out(sf, st);
    cudaFreeAsync(d_tab, st);
    if (d_carry) cudaFreeAsync(d_carry, st);
    const bool sync = sf.host != nullptr;
    release(sx, st);
    release(sf, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

int sgs_feat_stack(const double* feat, int n_sessions, int n_windows, int n_channels, int n_rows, int first_row,
                   int order, int step, double* out, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(n_sessions >= 1 && n_windows >= 0 && n_channels >= 1 && order >= 0 && step >= 1, "bad shape");
    if (n_rows <= 0) return SGS_OK;
    SGS_ARG(feat && out, "NULL argument");
    SGS_ARG(first_row >= 0 && first_row + n_rows <= n_windows, "rows [%d, %d) outside the %d windows", first_row, first_row + n_rows, n_windows);
    Staged sf, so;
    int rc = stage_in(sf, feat, sizeof(double) * (size_t)n_sessions * n_windows * n_channels, st);
    if (rc) return rc;
    rc = stage_out(so, out, sizeof(double) * (size_t)n_sessions * n_rows * n_channels * (order + 1), st);
    if (rc) { release(sf, st); return rc; }
    rc = stack_run((const double*)sf.dev, (double*)so.dev, n_sessions, n_windows, n_channels, n_rows, first_row, order, step, st);
    if (rc == SGS_OK) rc = finish_out(so, st);
    const bool sync = so.host != nullptr;
    release(sf, st);
    release(so, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// streaming feature extraction (ECogFeatCalc node)
// ---------------------------------------------------------------------------------------------------------------

struct sgs_feat_stream {
    sgs_feat_plan* plan = nullptr;
    int n_channels = 0, frame_size = 0, order = 0, step = 0;
    long long consumed = 0;
    bool cold = false;                  // ECogFeatCalc(warm_start=False): the last filter starts cold, no zero fill in front
    double *d_z = nullptr, *d_sq = nullptr, *d_feat = nullptr, *d_out = nullptr;
    void* d_x = nullptr;
    size_t x_cap = 0;
};

namespace sgs {
int feat_stream_row_width(const sgs_feat_stream* s) { return s->n_channels * (s->order + 1); }

/* Everything of a push except the read-back: copy the new samples (host or device) into the stream's staging
 * buffer and launch k_feat_stream; the stacked rows land in d_rows (or the stream's own buffer when NULL). */
int feat_stream_enqueue(sgs_feat_stream* s, const void* x, int x_is_f64, int n, const int64_t* frame_ends,
                        const int64_t* frame_index, int n_frames, double* d_rows, cudaStream_t st, bool src_pinned) {
    SGS_ARG(s && x && n >= 1, "bad arguments");
    SGS_ARG(n_frames >= 0 && n_frames <= kMaxFramesPerPush && (n_frames == 0 || (frame_ends && frame_index)), "bad frame schedule");
    SGS_ARG(n <= 128, "push at most 128 samples per call (got %d)", n);
    const size_t esz = x_is_f64 ? 8 : 4, bytes = (size_t)n * s->n_channels * esz;
    if (s->x_cap < bytes) {
        if (s->d_x) SGS_CUDA(cudaFree(s->d_x));
        SGS_CUDA(cudaMalloc(&s->d_x, bytes * 2));
        s->x_cap = bytes * 2;
    }
    StreamFrames fr;
    memset(&fr, 0, sizeof(fr));
    fr.n = n_frames;
    for (int q = 0; q < n_frames; ++q) {
        SGS_ARG(frame_ends[q] <= s->consumed + n && frame_ends[q] > s->consumed - 128, "frame %d does not end inside this push", q);
        fr.end[q] = frame_ends[q];
        fr.index[q] = frame_index[q];
    }
    int rc = copy_in_small(s->d_x, x, bytes, src_pinned, st);
    if (rc != SGS_OK) return rc;
    rc = feat_stream_run(s->plan->n_biquads, s->d_x, x_is_f64 != 0, n, s->n_channels, s->consumed, s->d_z, s->d_sq, s->d_feat,
                             s->plan->d_zf, s->cold ? 0 : s->plan->zero_fill, s->cold ? 1 : 0, s->frame_size, s->order, s->step,
                             d_rows ? d_rows : s->d_out, s->plan->cf, fr, st);
    if (rc != SGS_OK) return rc;
    s->consumed += n;
    return SGS_OK;
}
}  // namespace sgs

extern "C" {

void sgs_feat_stream_destroy(sgs_feat_stream* s) {
    if (!s) return;
    cudaFree(s->d_z); cudaFree(s->d_sq); cudaFree(s->d_feat); cudaFree(s->d_out); cudaFree(s->d_x);
    delete s;
}

int sgs_feat_stream_set_cold_start(sgs_feat_stream* s, int cold) {
    SGS_ARG(s, "NULL argument");
    SGS_ARG(s->consumed == 0, "the start mode can only be chosen before the first push");
    s->cold = cold != 0;
    return SGS_OK;
}

int sgs_feat_stream_create(sgs_feat_stream** stream_out, sgs_feat_plan* plan, int n_channels, int frame_size, int order, int step) {
    using namespace sgs;
    SGS_ARG(stream_out && plan && n_channels >= 1, "bad arguments");
    SGS_ARG(frame_size >= 1 && frame_size + 128 <= kSqRing, "frame_size %d too long for the streaming ring", frame_size);
    SGS_ARG(order >= 0 && step >= 1 && order * step + 1 <= kFeatRing, "context %d x %d too long", order, step);
    sgs_feat_stream* s = new sgs_feat_stream();
    s->plan = plan; s->n_channels = n_channels; s->frame_size = frame_size; s->order = order; s->step = step;
    const size_t C = n_channels;
    cudaError_t e = cudaMalloc((void**)&s->d_z, sizeof(double) * 2 * plan->n_biquads * C);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_sq, sizeof(double) * kSqRing * C);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_feat, sizeof(double) * kFeatRing * C);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_out, sizeof(double) * kMaxFramesPerPush * C * (order + 1));
    if (e == cudaSuccess) e = cudaMemset(s->d_feat, 0, sizeof(double) * kFeatRing * C);
    if (e != cudaSuccess) { sgs_feat_stream_destroy(s); return cuda_fail(e, "stream state", __FILE__, __LINE__); }
    *stream_out = s;
    return SGS_OK;
}

/* x: n x n_channels new samples (host).  frame_ends[n_frames] = exclusive end positions, in real-sample coordinates,
 * of the frames this push completes (host schedule, FrameBuffer.py:177); frame_index[n_frames] their running numbers.
 * out: n_frames x (n_channels*(order+1)) stacked rows (host). */
int sgs_feat_stream_push(sgs_feat_stream* s, const void* x, int x_is_f64, int n, const int64_t* frame_ends,
                         const int64_t* frame_index, int n_frames, double* out, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(n_frames == 0 || out, "bad frame schedule");
    int rc = feat_stream_enqueue(s, x, x_is_f64, n, frame_ends, frame_index, n_frames, nullptr, st);
    if (rc != SGS_OK) return rc;
    if (n_frames > 0) {
        SGS_CUDA(cudaMemcpyAsync(out, s->d_out, sizeof(double) * (size_t)n_frames * s->n_channels * (s->order + 1), cudaMemcpyDeviceToHost, st));
        SGS_CUDA(cudaStreamSynchronize(st));
    }
    return SGS_OK;
}

/* ---- filtering FrameBuffer node (livenodes/FrameBuffer.py:86-143): one SOS cascade, state resident between chunks ---- */
struct sgs_sos_stream {
    sgs::SosCoefs cf;
    int n_channels = 0, warm_start = 0;
    bool first = true;
    double *d_z = nullptr, *d_y = nullptr;
    void* d_x = nullptr;
    size_t cap = 0;                         // samples the staging buffers hold
};

void sgs_sos_stream_destroy(sgs_sos_stream* s) {
    if (!s) return;
    cudaFree(s->d_z); cudaFree(s->d_y); cudaFree(s->d_x);
    delete s;
}

int sgs_sos_stream_create(sgs_sos_stream** stream_out, const double* sos, const double* zi, int n_sections, int n_channels,
                          int warm_start) {
    using namespace sgs;
    SGS_ARG(stream_out && sos && zi, "NULL argument");
    SGS_ARG(n_sections >= 1 && n_sections <= kSosMaxSections && n_channels >= 1, "1..%d sections, >= 1 channel", kSosMaxSections);
    sgs_sos_stream* s = new sgs_sos_stream();
    memset(&s->cf, 0, sizeof(s->cf));
    s->cf.n_sections = n_sections; s->n_channels = n_channels; s->warm_start = warm_start ? 1 : 0;
    for (int k = 0; k < n_sections; ++k) {
        if (sos[k * 6 + 3] != 1.0) { delete s; set_error("sos section %d is not normalised (a0 = %g)", k, sos[k * 6 + 3]); return SGS_ERR_ARG; }
        s->cf.c[k][0] = sos[k * 6 + 0]; s->cf.c[k][1] = sos[k * 6 + 1]; s->cf.c[k][2] = sos[k * 6 + 2];
        s->cf.c[k][3] = sos[k * 6 + 4]; s->cf.c[k][4] = sos[k * 6 + 5];
        s->cf.zi[k][0] = zi[k * 2]; s->cf.zi[k][1] = zi[k * 2 + 1];
    }
    cudaError_t e = cudaMalloc((void**)&s->d_z, sizeof(double) * 2 * n_sections * n_channels);
    if (e != cudaSuccess) { sgs_sos_stream_destroy(s); return cuda_fail(e, "filter state", __FILE__, __LINE__); }
    *stream_out = s;
    return SGS_OK;
}

/* x[n][n_channels] (host, float32 or float64) -> y[n][n_channels] float64 (host): sosfilt with the state carried from the
 * previous push; the first push starts from zi (warm start) or zi * x[0] (cold start).  Synchronous. */
int sgs_sos_stream_push(sgs_sos_stream* s, const void* x, int x_is_f64, int n, double* y, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(s && x && y && n >= 1, "bad arguments");
    const size_t count = (size_t)n * s->n_channels;
    if (s->cap < count) {
        SGS_CUDA(cudaStreamSynchronize(st));
        cudaFree(s->d_x); cudaFree(s->d_y);
        s->d_x = nullptr; s->d_y = nullptr; s->cap = 0;
        SGS_CUDA(cudaMalloc(&s->d_x, count * 8 * 2));
        SGS_CUDA(cudaMalloc((void**)&s->d_y, count * 8 * 2));
        s->cap = count * 2;
    }
    SGS_CUDA(cudaMemcpyAsync(s->d_x, x, count * (x_is_f64 ? 8 : 4), cudaMemcpyDefault, st));
    int rc = sos_stream_run(s->d_x, x_is_f64 != 0, n, s->n_channels, s->d_z, s->d_y, s->first ? 1 : 0, s->warm_start, s->cf, st);
    if (rc != SGS_OK) return rc;
    s->first = false;
    SGS_CUDA(cudaMemcpyAsync(y, s->d_y, count * 8, cudaMemcpyDefault, st));
    SGS_CUDA(cudaStreamSynchronize(st));
    return SGS_OK;
}

}  // extern "C"
