// C ABI for LDA decoding (include/sgs.h).
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "kernels.cuh"
#include "../../include/sgs.h"


// tensor-core path geometry (csrc/lda_tc.cu)
static const int kTcN = 128, kTcK = 160, kTcClasses = 9, kTcSlices = 3;
static const unsigned kTcLBO = 2048, kTcSBO = 128;
// |tensor-core score - exact| <= eps |x'|_2 |w|_2: operand split 3 * 2^-24 = 1.8e-7 per product (hi + lo keep 24 bits each, the
// lo.lo term is dropped) + fp32 accumulation, worst case one 2^-23 truncation of the running sum per MMA over 60 MMAs = 7.2e-6
static const double kTcEps = 1e-5;
static const long long kTcMinFrames = 4096; // below this the fp64 kernel alone is faster than the extra launches

static float tf32_round(float x) {           // round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32.f32
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

struct sgs_lda_model {
    int n_bins = 0, n_classes = 0, n_features = 0, n_levels = 0, smooth_radius = 0, max_col = 0;
    double *d_Wt = nullptr, *d_bias = nullptr, *d_cls = nullptr, *d_medians = nullptr, *d_taps = nullptr;
    int* d_select = nullptr;
    // tensor-core path (built when the model fits it: 9 classes, <= 160 features, <= 42 bins)
    bool tc_ok = false;
    float* d_Bmat = nullptr;                  // [3][hi, lo][canonical 128 x 160 tf32]
    double *d_cls_tc = nullptr, *d_bias_tc = nullptr, *d_centre = nullptr, *d_chan_mean = nullptr;
    int *d_slice_bins = nullptr, *d_feat_chan = nullptr, *d_feat_back = nullptr, *d_iota = nullptr, *d_count = nullptr;
    int iota_len = 0, tab_order = -1, tab_step = -1;
    double* d_wnorm = nullptr;               // [3][16] largest class-weight norm of each bin
    std::vector<int32_t> h_select;
};

extern "C" {

void sgs_lda_model_destroy(sgs_lda_model* m) {
    if (!m) return;
    cudaFree(m->d_Wt); cudaFree(m->d_bias); cudaFree(m->d_cls); cudaFree(m->d_medians); cudaFree(m->d_taps); cudaFree(m->d_select);
    cudaFree(m->d_Bmat); cudaFree(m->d_cls_tc); cudaFree(m->d_bias_tc); cudaFree(m->d_centre); cudaFree(m->d_chan_mean);
    cudaFree(m->d_slice_bins); cudaFree(m->d_feat_chan); cudaFree(m->d_feat_back); cudaFree(m->d_iota); cudaFree(m->d_count); cudaFree(m->d_wnorm);
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
    m->h_select.assign(select, select + n_features);
    // ---- tensor-core tables: W split into tf32 hi + lo, slice-wise, in the canonical K-major core-matrix layout ----
    if (e == cudaSuccess && n_classes == kTcClasses && n_features <= kTcK && n_bins <= kTcSlices * (kTcN / kTcClasses)) {
        const int per = (n_bins + kTcSlices - 1) / kTcSlices;
        int slice_bins[kTcSlices + 1];
        for (int s = 0; s <= kTcSlices; ++s) slice_bins[s] = std::min(n_bins, s * per);
        const size_t mat = (size_t)(kTcK / 4) * kTcLBO / 4;              // floats per 128 x 160 operand matrix
        std::vector<float> B((size_t)kTcSlices * 2 * mat, 0.0f);
        std::vector<double> cls_tc((size_t)kTcSlices * kTcN, 0.0);
        std::vector<double> wn((size_t)kTcSlices * 16, 0.0);
        for (int s = 0; s < kTcSlices; ++s)
            for (int b = slice_bins[s]; b < slice_bins[s + 1]; ++b)
                for (int k = 0; k < n_classes; ++k) {
                    const int n = (b - slice_bins[s]) * kTcClasses + k;
                    cls_tc[(size_t)s * kTcN + n] = class_labels[b * n_classes + k];
                    double nrm = 0.0;
                    for (int f = 0; f < n_features; ++f) {
                        const double w = W[((size_t)b * n_classes + k) * n_features + f];
                        nrm += w * w;
                        const float hi = tf32_round((float)w);
                        const float lo = tf32_round((float)(w - (double)hi));
                        const size_t off = ((size_t)(f / 4) * kTcLBO + (size_t)(n / 8) * kTcSBO + (n % 8) * 16 + (f % 4) * 4) / 4;
                        B[((size_t)s * 2 + 0) * mat + off] = hi;
                        B[((size_t)s * 2 + 1) * mat + off] = lo;
                    }
                    wn[(size_t)s * 16 + (b - slice_bins[s])] = std::max(wn[(size_t)s * 16 + (b - slice_bins[s])], sqrt(nrm));
                }
        if (e == cudaSuccess) e = up((void**)&m->d_wnorm, wn.data(), wn.size() * sizeof(double));
        if (e == cudaSuccess) e = up((void**)&m->d_Bmat, B.data(), B.size() * sizeof(float));
        if (e == cudaSuccess) e = up((void**)&m->d_cls_tc, cls_tc.data(), cls_tc.size() * sizeof(double));
        if (e == cudaSuccess) e = up((void**)&m->d_slice_bins, slice_bins, sizeof(slice_bins));
        if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_bias_tc, sizeof(double) * kTcSlices * kTcN);
        if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_centre, sizeof(double) * kTcK);
        if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_feat_chan, sizeof(int) * kTcK);
        if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_feat_back, sizeof(int) * kTcK);
        if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_count, 2 * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(m->d_count, 0, 2 * sizeof(int));
        m->tc_ok = (e == cudaSuccess);
    }
    if (e != cudaSuccess) { sgs_lda_model_destroy(m); return sgs::cuda_fail(e, "model upload", __FILE__, __LINE__); }
    *model = m;
    return SGS_OK;
}

int sgs_lda_decode(const sgs_lda_model* cm, const double* feat, int n_sessions, int n_windows, int n_channels, int n_rows,
                   int first_row, int order, int step, double* labels, double* spec, int smooth, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    sgs_lda_model* m = const_cast<sgs_lda_model*>(cm);           // per-call scratch tables live in the handle
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
    const long long n_frames = (long long)n_sessions * n_rows;
    const char* env_tc = getenv("SGS_LDA_TC");
    const bool use_tc = m->tc_ok && n_frames >= kTcMinFrames && !(env_tc && env_tc[0] == '0');
    if (rc == SGS_OK && !use_tc)
        rc = lda_run((const double*)sf.dev, m->d_Wt, m->d_bias, m->d_cls, m->d_select, m->d_medians, m->d_taps,
                     (double*)sl.dev, (double*)ss.dev, smooth, n_sessions, g, st, nullptr, nullptr, 0);
    if (rc == SGS_OK && use_tc) {
        // tensor-core scoring as a filter + exact fp64 re-scoring of the flagged frames (csrc/lda_tc.cu)
        cudaError_t e = cudaSuccess;
        if (m->tab_order != order || m->tab_step != step) {
            std::vector<int> chan(kTcK, 0), back(kTcK, 0);
            for (int f = 0; f < m->n_features; ++f) {
                const int col = m->h_select[f];
                chan[f] = col / (order + 1);
                back[f] = (order - (col - chan[f] * (order + 1))) * step;
            }
            e = cudaMemcpyAsync(m->d_feat_chan, chan.data(), sizeof(int) * kTcK, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_feat_back, back.data(), sizeof(int) * kTcK, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);                     // the vectors go out of scope
            m->tab_order = order; m->tab_step = step;
        }
        if (e == cudaSuccess && m->iota_len < n_channels) {
            cudaFree(m->d_iota); cudaFree(m->d_chan_mean);
            std::vector<int> iota(n_channels);
            for (int i = 0; i < n_channels; ++i) iota[i] = i;
            e = cudaMalloc((void**)&m->d_iota, sizeof(int) * n_channels);
            if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_chan_mean, sizeof(double) * n_channels);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_iota, iota.data(), sizeof(int) * n_channels, cudaMemcpyHostToDevice);
            m->iota_len = n_channels;
        }
        double* d_lab = (double*)sl.dev;
        int *d_flags = nullptr, *d_list = nullptr, *d_count = m->d_count;
        bool own_lab = false;
        if (e == cudaSuccess && !d_lab) { e = cudaMallocAsync((void**)&d_lab, out_bytes, st); own_lab = true; }
        // one bin mask per (frame, slice of <= 14 bins); the list holds the non-zero ones
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_flags, sizeof(int) * 3 * n_frames, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_list, sizeof(int) * 3 * n_frames, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "tensor-core scratch", __FILE__, __LINE__);
        if (rc == SGS_OK)
            rc = col_means_run((const double*)sf.dev, (long long)n_sessions * n_windows, n_channels, m->d_iota, n_channels, m->d_chan_mean, st);
        if (rc == SGS_OK) {
            LdaTcGeom tg;
            tg.n_windows = n_windows; tg.n_channels = n_channels; tg.n_rows = n_rows; tg.first_row = first_row; tg.order = order;
            tg.step = step; tg.n_bins = m->n_bins; tg.n_features = m->n_features;
            tg.tiles_per_session = (n_rows + 127) / 128; tg.n_tiles = tg.tiles_per_session * n_sessions; tg.eps = kTcEps;
            rc = lda_tc_run((const double*)sf.dev, m->d_Bmat, m->d_Wt, m->d_bias, m->d_chan_mean, m->d_bias_tc, m->d_cls_tc, m->d_feat_chan,
                            m->d_feat_back, m->d_centre, m->d_slice_bins, m->d_wnorm, d_lab, d_flags, d_list, d_count, n_frames, tg, st);
        }
        if (rc == SGS_OK)
            rc = lda_pairs_run((const double*)sf.dev, m->d_Wt, m->d_bias, m->d_cls, m->d_select, d_lab, g, st, d_flags, d_list, d_count,
                               m->d_slice_bins, 3 * n_frames);
        if (rc == SGS_OK && spec)
            rc = dequantize_run(d_lab, m->d_medians, m->d_taps, m->smooth_radius, smooth, m->n_bins, m->n_levels, n_frames, (double*)ss.dev, st);
        if (own_lab && d_lab) cudaFreeAsync(d_lab, st);
        if (d_flags) cudaFreeAsync(d_flags, st);
        if (d_list) cudaFreeAsync(d_list, st);
    }
    if (rc == SGS_OK && labels) rc = finish_out(sl, st);
    if (rc == SGS_OK && spec) rc = finish_out(ss, st);
    const bool sync = sl.host || ss.host;
    release(sf, st); release(sl, st); release(ss, st);
    if (rc == SGS_OK && sync) SGS_CUDA(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"

namespace sgs {
int lda_model_bins(const sgs_lda_model* m) { return m->n_bins; }

/* Score n_rows already-stacked feature rows resident on the device (the streaming chain): labels and the
 * dequantised (optionally smoothed) spectrum, fp64 kernel, nothing staged, nothing synchronised. */
int lda_rows_enqueue(const sgs_lda_model* m, const double* d_rows, int n_rows, int row_width, double* d_labels, double* d_spec,
                     int smooth, cudaStream_t st) {
    SGS_ARG(m && d_rows && (d_labels || d_spec) && n_rows >= 1, "bad arguments");
    SGS_ARG(m->max_col < row_width, "select refers to column %d but frames have %d", m->max_col, row_width);
    SGS_ARG(!smooth || m->smooth_radius > 0, "model was created without smoothing taps");
    LdaGeom g;
    g.n_bins = m->n_bins; g.n_classes = m->n_classes; g.n_features = m->n_features; g.n_levels = m->n_levels;
    g.n_windows = n_rows; g.n_channels = row_width; g.n_rows = n_rows; g.first_row = 0; g.order = 0; g.step = 1;
    g.smooth_radius = m->smooth_radius;
    return lda_run(d_rows, m->d_Wt, m->d_bias, m->d_cls, m->d_select, m->d_medians, m->d_taps, d_labels, d_spec, smooth, 1, g, st,
                   nullptr, nullptr, 0);
}
}  // namespace sgs

extern "C" {

/* Diagnostics: number of frames the last tensor-core decode handed to the exact fp64 re-scoring pass. */
int sgs_lda_last_rescored(const sgs_lda_model* m, int* n_frames) {
    SGS_ARG(m && n_frames, "NULL argument");
    *n_frames = 0;
    if (!m->d_count) return SGS_OK;
    SGS_CUDA(cudaDeviceSynchronize());
    SGS_CUDA(cudaMemcpy(n_frames, m->d_count + 1, sizeof(int), cudaMemcpyDeviceToHost));
    return SGS_OK;
}

}  // extern "C"
