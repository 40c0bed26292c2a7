// C ABI for the fused streaming chain (include/sgs.h): one packet of sEEG samples in, every product of the
// ECogFeatCalc -> LDASynthesis -> Dequantization -> GriffinLimSynthesis node chain out, with ONE host<->device
// round trip per packet instead of one per node and frame.
//
// Reference call chain replaced: decode.py:152-183 wires the four nodes; each 10 ms frame then walks
// ECogFeatCalc.py:67-104 -> LDASynthesis.py:19-28 -> Dequantization.py:15-18 -> GriffinLim.py:98-174 synchronously.
// Here the same four device streams (their state is unchanged and shared with the single-node entry points) are
// driven back to back on one CUDA stream; inputs and results go through page-locked staging owned by the chain.
#include <algorithm>
#include "kernels.cuh"
#include "../../include/sgs.h"

namespace sgs {
constexpr int kChainMaxFrames = 16, kChainMaxSamples = 128, kChainBlk = 480, kChainHopMax = 192;
}  // namespace sgs

struct sgs_chain {
    sgs_feat_stream* feat = nullptr;
    const sgs_lda_model* lda = nullptr;
    sgs_gl_node* gl = nullptr;
    int row_width = 0, n_bins = 0, n_channels = 0;
    // page-locked staging: [samples | noise] in, [rows | labels | spectrum | pcm] out; device mirror of the outputs
    char *h_in = nullptr, *h_out = nullptr, *d_out = nullptr;
    size_t in_x = 0, in_noise = 0, out_bytes = 0;
    cudaEvent_t in_consumed = nullptr;        // the staging of the previous push has been read by the device
};

extern "C" {

void sgs_chain_destroy(sgs_chain* c) {
    if (!c) return;
    cudaFreeHost(c->h_in); cudaFreeHost(c->h_out); cudaFree(c->d_out);
    if (c->in_consumed) cudaEventDestroy(c->in_consumed);
    delete c;
}

int sgs_chain_create(sgs_chain** chain, sgs_feat_stream* feat, int n_channels, const sgs_lda_model* lda, sgs_gl_node* gl) {
    using namespace sgs;
    SGS_ARG(chain && feat && lda && gl && n_channels >= 1, "NULL argument");
    sgs_chain* c = new sgs_chain();
    c->feat = feat; c->lda = lda; c->gl = gl; c->n_channels = n_channels;
    c->row_width = feat_stream_row_width(feat);
    c->n_bins = lda_model_bins(lda);
    c->in_x = 0;
    c->in_noise = sizeof(double) * kChainMaxSamples * n_channels;
    const size_t in_bytes = c->in_noise + sizeof(double) * kChainMaxFrames * kChainBlk;
    c->out_bytes = sizeof(double) * kChainMaxFrames * (c->row_width + 2 * c->n_bins) + sizeof(short) * kChainMaxFrames * kChainHopMax;
    cudaError_t e = cudaHostAlloc((void**)&c->h_in, in_bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_out, c->out_bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_out, c->out_bytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->in_consumed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(c->in_consumed, 0);
    if (e != cudaSuccess) { sgs_chain_destroy(c); return cuda_fail(e, "chain staging", __FILE__, __LINE__); }
    *chain = c;
    return SGS_OK;
}

int sgs_chain_push(sgs_chain* c, const void* x, int x_is_f64, int n, const int64_t* frame_ends, const int64_t* frame_index,
                   int n_frames, const int32_t* gl_pos, int32_t gl_pos_before, const double* noise, uint64_t seed,
                   double* rows, double* labels, double* spec, int16_t* pcm, int* n_pcm, void* stream) {
    using namespace sgs;
    cudaStream_t st = (cudaStream_t)stream;
    SGS_ARG(c && x && n >= 1 && n <= kChainMaxSamples, "push takes 1..%d samples (got %d)", kChainMaxSamples, n);
    SGS_ARG(n_frames >= 0 && n_frames <= kChainMaxFrames, "push completes at most %d frames (got %d)", kChainMaxFrames, n_frames);
    SGS_ARG(n_frames == 0 || (gl_pos && rows && labels && spec && pcm && n_pcm), "NULL output");
    SGS_ARG(!is_device_ptr(x), "sgs_chain_push takes host samples");
    const size_t xbytes = (size_t)n * c->n_channels * (x_is_f64 ? 8 : 4);
    SGS_CUDA(cudaEventSynchronize(c->in_consumed));   // a frame-less push returns without synchronising
    memcpy(c->h_in + c->in_x, x, xbytes);
    if (noise && n_frames > 0) memcpy(c->h_in + c->in_noise, noise, sizeof(double) * (size_t)n_frames * kChainBlk);
    // outputs packed for this push's frame count, so the read-back moves only what was produced
    const size_t o_rows = 0, o_labels = o_rows + sizeof(double) * (size_t)n_frames * c->row_width,
                 o_spec = o_labels + sizeof(double) * (size_t)n_frames * c->n_bins,
                 o_pcm = o_spec + sizeof(double) * (size_t)n_frames * c->n_bins;
    double* d_rows = (double*)(c->d_out + o_rows);
    double* d_labels = (double*)(c->d_out + o_labels);
    double* d_spec = (double*)(c->d_out + o_spec);
    short* d_pcm = (short*)(c->d_out + o_pcm);
    int rc = feat_stream_enqueue(c->feat, c->h_in + c->in_x, x_is_f64, n, frame_ends, frame_index, n_frames, d_rows, st, true);
    if (rc != SGS_OK) return rc;
    if (n_pcm) *n_pcm = 0;
    if (n_frames == 0) { SGS_CUDA(cudaEventRecord(c->in_consumed, st)); return SGS_OK; }                 // nothing to read back: the packet only advanced the filter state
    rc = lda_rows_enqueue(c->lda, d_rows, n_frames, c->row_width, d_labels, d_spec, 1, st);
    if (rc != SGS_OK) return rc;
    int total = 0;
    rc = gl_node_enqueue(c->gl, d_spec, n_frames, gl_pos, gl_pos_before, noise ? (const double*)(c->h_in + c->in_noise) : nullptr,
                         seed, d_pcm, &total, st, true);
    if (rc != SGS_OK) return rc;
    // one read-back of everything the four nodes emit for this packet
    const size_t used = o_pcm + sizeof(short) * (size_t)total;
    SGS_CUDA(cudaMemcpyAsync(c->h_out, c->d_out, used, cudaMemcpyDeviceToHost, st));
    SGS_CUDA(cudaStreamSynchronize(st));
    memcpy(rows, c->h_out + o_rows, o_labels - o_rows);
    memcpy(labels, c->h_out + o_labels, o_spec - o_labels);
    memcpy(spec, c->h_out + o_spec, o_pcm - o_spec);
    memcpy(pcm, c->h_out + o_pcm, sizeof(short) * (size_t)total);
    *n_pcm = total;
    return SGS_OK;
}

}  // extern "C"
