// C ABI for LDA decoding (include/sgs.h).
#include <math.h>
#include <vector>
#include "common.cuh"
#include "../../include/sgs.h"

namespace sgs {
struct LdaGeom {
    int n_bins, n_classes, n_features, n_levels;
    int n_windows, n_channels, n_rows, first_row, order, step;
    int smooth_radius;
};
int lda_run(const double* feat, const double* Wt, const double* bias, const double* cls, const int* select,
            const double* medians, const double* taps, double* labels, double* spec, int smooth, int n_sessions,
            const LdaGeom& g, cudaStream_t st);
}  // namespace sgs

struct sgs_lda_model {
    int n_bins = 0, n_classes = 0, n_features = 0, n_levels = 0, smooth_radius = 0, max_col = 0;
    double *d_Wt = nullptr, *d_bias = nullptr, *d_cls = nullptr, *d_medians = nullptr, *d_taps = nullptr;
    int* d_select = nullptr;
};

extern "C" {

void sgs_lda_model_destroy(sgs_lda_model* m) {
    if (!m) return;
    cudaFree(m->d_Wt); cudaFree(m->d_bias); cudaFree(m->d_cls); cudaFree(m->d_medians); cudaFree(m->d_taps); cudaFree(m->d_select);
    delete m;
}

int sgs_lda_model_create(sgs_lda_model** model, int n_bins, int n_classes, int n_features, const double* W,
                         const double* bias, const double* class_labels, const int32_t* select, const double* medians,
                         int n_levels, const double* smooth_taps, int smooth_radius) {
    SGS_ARG(model && W && bias && class_labels && select && medians, "NULL argument");
    SGS_ARG(n_bins >= 1 && n_classes >= 1 && n_features >= 1 && n_levels >= 1, "bad model shape");
    SGS_ARG(smooth_radius >= 0 && (smooth_radius == 0 || smooth_taps), "smoothing taps missing");
    SGS_ARG(smooth_radius < n_bins, "smoothing radius %d >= number of bins %d", smooth_radius, n_bins);
    sgs_lda_model* m = new sgs_lda_model();
    m->n_bins = n_bins; m->n_classes = n_classes; m->n_features = n_features; m->n_levels = n_levels;
    m->smooth_radius = smooth_radius;
    for (int f = 0; f < n_features; ++f) {
        if (select[f] < 0) { delete m; sgs::set_error("negative feature index in select"); return SGS_ERR_ARG; }
        if (select[f] > m->max_col) m->max_col = select[f];
    }
    // repack [bin][class][feature] -> [bin][feature][class]
    std::vector<double> wt((size_t)n_bins * n_features * n_classes);
    for (int b = 0; b < n_bins; ++b)
        for (int k = 0; k < n_classes; ++k)
            for (int f = 0; f < n_features; ++f)
                wt[((size_t)b * n_features + f) * n_classes + k] = W[((size_t)b * n_classes + k) * n_features + f];
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
        return e;
    };
    cudaError_t e = up((void**)&m->d_Wt, wt.data(), wt.size() * sizeof(double));
    if (e == cudaSuccess) e = up((void**)&m->d_bias, bias, sizeof(double) * n_bins * n_classes);
    if (e == cudaSuccess) e = up((void**)&m->d_cls, class_labels, sizeof(double) * n_bins * n_classes);
    if (e == cudaSuccess) e = up((void**)&m->d_medians, medians, sizeof(double) * n_bins * n_levels);
    if (e == cudaSuccess) e = up((void**)&m->d_select, select, sizeof(int32_t) * n_features);
    if (e == cudaSuccess && smooth_radius > 0) e = up((void**)&m->d_taps, smooth_taps, sizeof(double) * (2 * smooth_radius + 1));
    if (e != cudaSuccess) { sgs_lda_model_destroy(m); return sgs::cuda_fail(e, "model upload", __FILE__, __LINE__); }
    *model = m;
    return SGS_OK;
}

int sgs_lda_decode(const sgs_lda_model* m, const double* feat, int n_sessions, int n_windows, int n_channels, int n_rows,
                   int first_row, int order, int step, double* labels, double* spec, int smooth, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(m != nullptr, "model is NULL");
    SGS_ARG(n_sessions >= 1 && n_windows >= 0 && n_channels >= 1 && order >= 0 && step >= 1, "bad shape");
    if (n_rows <= 0) return SGS_OK;
    SGS_ARG(feat != nullptr && (labels || spec), "NULL argument");
    SGS_ARG(first_row >= 0 && first_row + n_rows <= n_windows, "rows [%d, %d) outside the %d windows", first_row, first_row + n_rows, n_windows);
    SGS_ARG(m->max_col < n_channels * (order + 1), "select refers to column %d but frames have %d", m->max_col, n_channels * (order + 1));
    SGS_ARG(!smooth || m->smooth_radius > 0, "model was created without smoothing taps");
    LdaGeom g;
    g.n_bins = m->n_bins; g.n_classes = m->n_classes; g.n_features = m->n_features; g.n_levels = m->n_levels;
    g.n_windows = n_windows; g.n_channels = n_channels; g.n_rows = n_rows; g.first_row = first_row; g.order = order; g.step = step;
    g.smooth_radius = m->smooth_radius;
    const size_t out_bytes = sizeof(double) * (size_t)n_sessions * n_rows * m->n_bins;
    Staged sf, sl, ss;
    int rc = stage_in(sf, feat, sizeof(double) * (size_t)n_sessions * n_windows * n_channels, st);
    if (rc == SGS_OK && labels) rc = stage_out(sl, labels, out_bytes, st);
    if (rc == SGS_OK && spec) rc = stage_out(ss, spec, out_bytes, st);
    if (rc == SGS_OK)
        rc = lda_run((const double*)sf.dev, m->d_Wt, m->d_bias, m->d_cls, m->d_select, m->d_medians, m->d_taps,
                     (double*)sl.dev, (double*)ss.dev, smooth, n_sessions, g, st);
    if (rc == SGS_OK && labels) rc = finish_out(sl, st);
    if (rc == SGS_OK && spec) rc = finish_out(ss, st);
    const bool sync = sl.host || ss.host;
    release(sf, st); release(sl, st); release(ss, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"
