// C ABI for the training-side kernels (include/sgs.h).
#include <vector>
#include "kernels.cuh"
#include "../../include/sgs.h"


namespace {
struct Bufs {
    std::vector<sgs::Staged*> all;
    cudaStream_t st;
    int rc = SGS_OK;
    bool host_out = false;
    explicit Bufs(cudaStream_t s) : st(s) {}
    void* in(sgs::Staged& s, const void* p, size_t bytes) { if (rc == SGS_OK) rc = sgs::stage_in(s, p, bytes, st); all.push_back(&s); return s.dev; }
    void* out(sgs::Staged& s, void* p, size_t bytes) { if (rc == SGS_OK) rc = sgs::stage_out(s, p, bytes, st); all.push_back(&s); if (s.host) host_out = true; return s.dev; }
    int finish(std::initializer_list<sgs::Staged*> outs) {
        for (auto* o : outs) if (rc == SGS_OK) rc = sgs::finish_out(*o, st);
        for (auto* s : all) sgs::release(*s, st);
        if (rc == SGS_OK && host_out) { cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) rc = sgs::cuda_fail(e, "sync", __FILE__, __LINE__); }
        return rc;
    }
};
}  // namespace

extern "C" {

int sgs_col_minmax(const double* y, int64_t n, int ncol, double* mn, double* mx, void* stream) {
    using namespace sgs;
    SGS_ARG(y && mn && mx && n >= 1 && ncol >= 1, "bad arguments");
    Bufs b((cudaStream_t)stream);
    Staged sy, s0, s1;
    b.in(sy, y, sizeof(double) * (size_t)n * ncol); b.out(s0, mn, sizeof(double) * ncol); b.out(s1, mx, sizeof(double) * ncol);
    if (b.rc == SGS_OK) b.rc = colminmax_run((const double*)sy.dev, n, ncol, (double*)s0.dev, (double*)s1.dev, b.st);
    return b.finish({&s0, &s1});
}

int sgs_quantize(const double* y, int64_t n, int ncol, const double* borders, int n_intervals, double* labels, void* stream) {
    using namespace sgs;
    SGS_ARG(n >= 0 && ncol >= 1 && n_intervals >= 1, "bad arguments");
    if (n == 0) return SGS_OK;
    SGS_ARG(y && borders && labels, "NULL argument");
    Bufs b((cudaStream_t)stream);
    Staged sy, sb, sq;
    b.in(sy, y, sizeof(double) * (size_t)n * ncol); b.in(sb, borders, sizeof(double) * ncol * n_intervals);
    b.out(sq, labels, sizeof(double) * (size_t)n * ncol);
    if (b.rc == SGS_OK) b.rc = quantize_run((const double*)sy.dev, n, ncol, (const double*)sb.dev, n_intervals, (double*)sq.dev, b.st);
    return b.finish({&sq});
}

int sgs_spearman(const double* x, int64_t n, int ncol, int64_t row_stride, const double* y, int ny, double* rho,
                 double* colsum, void* stream) {
    using namespace sgs;
    SGS_ARG(x && y && rho && colsum && n >= 2 && ncol >= 1 && ny >= 1, "bad arguments");
    if (row_stride == 0) row_stride = ncol;
    Bufs b((cudaStream_t)stream);
    Staged sx, sy, sr, sc;
    b.in(sx, x, sizeof(double) * ((size_t)(n - 1) * row_stride + ncol)); b.in(sy, y, sizeof(double) * (size_t)n * ny);
    b.out(sr, rho, sizeof(double) * ncol); b.out(sc, colsum, sizeof(double) * ncol);
    if (b.rc == SGS_OK)
        b.rc = spearman_run((const double*)sx.dev, n, ncol, row_stride, (const double*)sy.dev, ny, (double*)sr.dev, (double*)sc.dev, b.st);
    return b.finish({&sr, &sc});
}

int sgs_col_means(const double* x, int64_t n, int64_t row_stride, const int32_t* select, int n_features, double* xbar, void* stream) {
    using namespace sgs;
    SGS_ARG(x && select && xbar && n >= 1 && n_features >= 1 && row_stride >= 1, "bad arguments");
    int max_col = 0;
    for (int f = 0; f < n_features; ++f) { SGS_ARG(select[f] >= 0, "negative column"); if (select[f] > max_col) max_col = select[f]; }
    SGS_ARG(max_col < row_stride, "select refers to column %d of %lld", max_col, (long long)row_stride);
    Bufs b((cudaStream_t)stream);
    Staged sx, ss, so;
    b.in(sx, x, sizeof(double) * (size_t)n * row_stride); b.in(ss, select, sizeof(int32_t) * n_features);
    b.out(so, xbar, sizeof(double) * n_features);
    if (b.rc == SGS_OK) b.rc = col_means_run((const double*)sx.dev, n, row_stride, (const int*)ss.dev, n_features, (double*)so.dev, b.st);
    return b.finish({&so});
}

int sgs_lda_stats(const double* x, int64_t n, int64_t row_stride, const int32_t* select, int n_features, const double* labels,
                  int n_bins, int n_classes, const double* xbar_in, double* xbar, double* G, double* class_sums, double* counts,
                  void* stream) {
    using namespace sgs;
    SGS_ARG(x && select && labels && xbar && G && class_sums && counts, "NULL argument");
    SGS_ARG(n >= 1 && n_features >= 1 && n_bins >= 1 && n_classes >= 1 && row_stride >= 1, "bad shape");
    SGS_ARG(n_features <= 1024, "at most 1024 model features");
    int max_col = 0;
    for (int f = 0; f < n_features; ++f) { SGS_ARG(select[f] >= 0, "negative column"); if (select[f] > max_col) max_col = select[f]; }
    SGS_ARG(max_col < row_stride, "select refers to column %d of %lld", max_col, (long long)row_stride);
    Bufs b((cudaStream_t)stream);
    Staged sx, ss, sl, sbi, sb, sg, sc, sn;
    b.in(sx, x, sizeof(double) * (size_t)n * row_stride); b.in(ss, select, sizeof(int32_t) * n_features);
    b.in(sl, labels, sizeof(double) * (size_t)n * n_bins);
    if (xbar_in) b.in(sbi, xbar_in, sizeof(double) * n_features);
    b.out(sb, xbar, sizeof(double) * n_features); b.out(sg, G, sizeof(double) * (size_t)n_features * n_features);
    b.out(sc, class_sums, sizeof(double) * (size_t)n_bins * n_classes * n_features); b.out(sn, counts, sizeof(double) * n_bins * n_classes);
    if (b.rc == SGS_OK)
        b.rc = lda_stats_run((const double*)sx.dev, n, row_stride, (const int*)ss.dev, n_features, (const double*)sl.dev, n_bins, n_classes,
                             (double*)sb.dev, (double*)sg.dev, (double*)sc.dev, (double*)sn.dev, (const double*)sbi.dev, b.st);
    return b.finish({&sb, &sg, &sc, &sn});
}

}  // extern "C"
