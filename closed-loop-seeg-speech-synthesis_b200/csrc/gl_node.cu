// Streaming-semantics Griffin-Lim (the GriffinLimSynthesis node), batched over frames and sessions.
//
// Reference semantics restated (livenodes/GriffinLim.py, closed form R3 of SURVEY.md 8a'):
//   * per frame k >= 1 an independent 480-sample block is synthesised from the log-mels of frames k-1, k and a
//     fresh np.random.rand(480) start (GriffinLim.py:76-96): `iters` times { X_o = rfft(blackman256 * x[o:o+256])
//     for o in {0,160} (range(0, 480-256, 160), quirk Q3); Z_o = S_o * exp(angle(X_o))  -- REAL, the reference has
//     no 1j in the exponent (quirk Q1); x = sum_o irfft(Z_o) * blackman256 placed at o; x[416:480] = 0 }.
//   * S = fromLogMels: exp(logmel) through the 2-tap inverse mel matrix (local/MelFilterBank.py:82-83).
//   * ring-buffer overlap-add of the un-windowed blocks normalised by the sum of blackman(480) segments, with the
//     reference's write-head positions int((ms/1000)*16000) (GriffinLim.py:115-166; hops of 159/161 samples, Q7),
//     division skipped where the window sum is exactly 0.
//   * order-5 low-pass (scipy.signal.lfilter, state carried across hops), clip, int16 truncation (GriffinLim.py:169-174).
//
// k_gl_blocks8 (gl_blocks8.cuh): 8 lanes per STFT frame, FFTs in registers, one shared-memory transposition per
//              transform, exp(angle) as polynomial pieces (exp_angle.cuh); FP64 throughout (the branch cut of angle()
//              makes the iteration chaotic under fp32 round-off).
// k_gl_ola:    gathers the <= 4 blocks covering each output sample in arrival order.
// k_lp_*:      the IIR low-pass as an exact chunked scan over the 5-state recurrence (zero-state chunk pass,
//              sequential 5x5 carry, final pass with clip + int16).
#include <math.h>
#include <stdlib.h>
#include "kernels.cuh"
#include "exp_angle.cuh"

namespace sgs {



__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long item, unsigned idx) {
    // counter-based generator for throughput runs (parity runs pass the reference's MT19937 draws explicitly)
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (item * 480ULL + idx + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// fromLogMels for one bin through the 2-tap inverse mel matrix (MelFilterBank.py:82-83), out of line (once per block)
__device__ __noinline__ double mel_magnitude(const double* lm, const int* inv_idx, const double* inv_w, int bin, int log_mels) {
    const double w0 = inv_w[bin * 2], w1 = inv_w[bin * 2 + 1];
    double v = 0.0;
    if (w0 != 0.0) v = (log_mels ? exp(lm[inv_idx[bin * 2]]) : lm[inv_idx[bin * 2]]) * w0;
    if (w1 != 0.0) v = fma(log_mels ? exp(lm[inv_idx[bin * 2 + 1]]) : lm[inv_idx[bin * 2 + 1]], w1, v);
    return (log_mels && !isfinite(v)) ? 0.0 : v;                            // MelFilterBank.makeNormal (fromLogMels only)
}

// out-of-line copy of the four-wide evaluation: one body in the instruction cache, called 4 times per iteration (inlined
// four-, eight- and sixteen-wide it measured 29.1 / 30.4 / 35.6 ms per 1.92 M blocks against 28.8 ms: spills)
struct Ea4 { double a, b, c, d; };
__device__ __noinline__ Ea4 exp_angle4_call(const double* s_tab, double im0, double re0, double im1, double re1, double im2,
                                            double re2, double im3, double re3) {
    const double im[4] = {im0, im1, im2, im3}, re[4] = {re0, re1, re2, re3};
    double out[4];
    exp_angle_n<4>(s_tab, im, re, out);
    return Ea4{out[0], out[1], out[2], out[3]};
}

}  // namespace sgs
#include "gl_blocks8.cuh"
namespace sgs {

// ------------------------------------------------------------------------------------------------
// overlap-add + window-sum normalisation, linear-time restatement of the node's ring buffers.
// pos[k] = write head after frame k (host table, the reference's float expression); frame k returns the
// pos[k]-pos[k-1] samples starting at absolute position pos[k] - 480.
// ------------------------------------------------------------------------------------------------
constexpr int kOlaFrames = 32;       // frames per CTA: one CTA per frame was launch-bound (1.9 M tiny CTAs per bench step)
__global__ void __launch_bounds__(256)
k_gl_ola(const double* __restrict__ blocks, const int* __restrict__ pos, const double* __restrict__ ola_window,
         double* __restrict__ v, int n_frames, int first_frame, long long out_per_sess) {
    __shared__ int s_pos[kOlaFrames + 5];                  // pos[k0 - 5 .. k0 + 31]
    __shared__ double s_win[kBlk];
    const int k0 = first_frame + blockIdx.x * kOlaFrames, sess = blockIdx.y;
    const int k1 = k0 + kOlaFrames < n_frames ? k0 + kOlaFrames : n_frames;
    for (int i = threadIdx.x; i < kOlaFrames + 5; i += blockDim.x) {
        const int k = k0 - 5 + i;
        s_pos[i] = k < 0 ? 0 : pos[k < n_frames ? k : n_frames - 1];
    }
    for (int i = threadIdx.x; i < kBlk; i += blockDim.x) s_win[i] = ola_window[i];
    __syncthreads();
    const int out0 = first_frame > 0 ? pos[first_frame - 1] : 0;
    const double* bs = blocks + (long long)sess * n_frames * kBlk;
    for (int k = k0; k < k1; ++k) {
        const int pk = s_pos[k - k0 + 5], prev = s_pos[k - k0 + 4];           // pos[k], pos[k - 1] (0 before the first frame)
        const int shifted = pk - prev;
        if ((int)threadIdx.x < shifted) {
            const int i = threadIdx.x;
            const int p = pk - kBlk + i;                                       // absolute position of this output sample
            double num = 0.0, den = 0.0;
#pragma unroll
            for (int d = 4; d >= 0; --d) {                                     // arrival order = ascending frame
                const int jj = k - d;
                if (jj >= first_frame) {
                    const int off = p - (s_pos[jj - k0 + 5] - kBlk);
                    if (off >= 0 && off < kBlk) {
                        num += bs[(long long)jj * kBlk + off];
                        den += s_win[off];
                    }
                }
            }
            v[(long long)sess * out_per_sess + (prev - out0) + i] = (den != 0.0) ? num / den : num;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// order-ORD IIR (direct form II transposed, scipy.signal.lfilter) along each session's output stream
// ------------------------------------------------------------------------------------------------

template <int ORD>
__device__ __forceinline__ double lp_step(double xin, double (&z)[ORD], const LpCoefs& c) {
    const double y = fma(c.b[0], xin, z[0]);
#pragma unroll
    for (int i = 0; i < ORD - 1; ++i) z[i] = fma(-c.a[i + 1], y, fma(c.b[i + 1], xin, z[i + 1]));
    z[ORD - 1] = fma(-c.a[ORD], y, c.b[ORD] * xin);
    return y;
}

// Low-pass passes: thread = one 2048-sample chunk (exact two-pass scan: zero-state chunk responses, sequential
// carry with Phi = M^2048, final pass).  A warp owns 32 consecutive chunks of one session and stages them through
// shared memory 64 samples at a time, so every global access is a fully coalesced 256 B row instead of 32 lanes
// striding 16 KB apart.
constexpr int kLpTile = 32, kLpChunk = 2048, kLpWarps = 2;

// 8-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills (samples past the end of the stream)
__device__ __forceinline__ void lp_cp_async8(double* dst, const double* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}

// Tiles are double-buffered: the copies of tile tt+1 (cp.async, no registers) are in flight while the warp runs the 32
// recurrence steps of tile tt - the staging used to be bound by one HBM round trip per 16 rows.
template <int ORD, bool APPLY>
__global__ void __launch_bounds__(kLpWarps * 32)
k_lp_pass(const double* __restrict__ v, const double* __restrict__ start_states, double* __restrict__ end_states,
          short* __restrict__ pcm, double* __restrict__ filtered, const __grid_constant__ LpCoefs c, double norm_div,
          long long n_out, int n_chunks, int n_groups /*per session*/) {
    __shared__ double sm[kLpWarps][2][32][kLpTile + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sess = blockIdx.y;
    const int group = blockIdx.x * kLpWarps + warp;
    if (group >= n_groups) return;
    const int ci0 = group * 32, ci = ci0 + lane;
    const bool live = ci < n_chunks;
    const double* ps = v + (long long)sess * n_out;
    double z[ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) z[i] = (APPLY && live) ? start_states[((long long)sess * n_chunks + ci) * kLpMaxOrd + i] : 0.0;
    const long long my_t0 = (long long)ci * kLpChunk;
    constexpr int kTiles = kLpChunk / kLpTile;
    // stage: row r of the tile = samples [tt*32, tt*32+32) of chunk ci0 + r; one 256 B row per warp copy
    auto stage = [&](int tt, int buf) {
#pragma unroll
        for (int row = 0; row < 32; ++row) {
            const long long t = (long long)(ci0 + row) * kLpChunk + tt * kLpTile + lane;
            const bool ok = t < n_out;
            lp_cp_async8(&sm[warp][buf][row][lane], ps + (ok ? t : 0), ok ? 8 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, 0);
    for (int tt = 0; tt < kTiles; ++tt) {
        const int buf = tt & 1;
        if (tt + 1 < kTiles) { stage(tt + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        double (*tile)[kLpTile + 1] = sm[warp][buf];
        const long long tbase = my_t0 + tt * kLpTile;
        if (live && tbase < n_out) {
            const int cnt = (n_out - tbase) < kLpTile ? (int)(n_out - tbase) : kLpTile;
            if (cnt == kLpTile) {
#pragma unroll 8
                for (int j = 0; j < kLpTile; ++j) {
                    const double y = lp_step<ORD>(tile[lane][j], z, c);
                    if (APPLY) tile[lane][j] = y;
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    const double y = lp_step<ORD>(tile[lane][j], z, c);
                    if (APPLY) tile[lane][j] = y;
                }
            }
        }
        __syncwarp();
        if (APPLY) {
#pragma unroll 8
            for (int row = 0; row < 32; ++row) {
                const long long t = (long long)(ci0 + row) * kLpChunk + tt * kLpTile + lane;
                if (t < n_out) {
                    const double y = tile[row][lane];
                    if (filtered) filtered[(long long)sess * n_out + t] = y;
                    double qv = y / norm_div;                       // np.clip(y / (normFactor * 1.01), -0.99, 0.99) * 32767
                    qv = qv < -0.99 ? -0.99 : (qv > 0.99 ? 0.99 : qv);
                    pcm[(long long)sess * n_out + t] = (short)(int)(qv * 32767.0);      // np.int16(): truncation toward zero
                }
            }
            __syncwarp();                                           // the tile is refilled by the copies of tile tt + 2
        }
    }
    if (!APPLY && live) {
        double* o = end_states + ((long long)sess * n_chunks + ci) * kLpMaxOrd;
#pragma unroll
        for (int i = 0; i < kLpMaxOrd; ++i) o[i] = i < ORD ? z[i] : 0.0;
    }
}

// carry: start[ci] = true state at the start of chunk ci (start[0] = zi); zi receives the final state.
// thread = session; e (zero-state chunk responses) and start are separate arrays so the loads pipeline.
template <int ORD>
__global__ void k_lp_carry(const double* __restrict__ e, double* __restrict__ start, const double* __restrict__ phi,
                           double* __restrict__ zi, int n_chunks, int n_sessions) {
    const int sess = blockIdx.x * blockDim.x + threadIdx.x;
    if (sess >= n_sessions) return;
    double s[ORD], P[ORD][ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) {
        s[i] = zi[sess * ORD + i];
#pragma unroll
        for (int k = 0; k < ORD; ++k) P[i][k] = phi[i * ORD + k];
    }
    const double* ep = e + (long long)sess * n_chunks * kLpMaxOrd;
    double* sp = start + (long long)sess * n_chunks * kLpMaxOrd;
#pragma unroll 4
    for (int ci = 0; ci < n_chunks; ++ci) {
        double nx[ORD];
#pragma unroll
        for (int i = 0; i < ORD; ++i) {
            sp[ci * kLpMaxOrd + i] = s[i];
            double acc = ep[ci * kLpMaxOrd + i];
#pragma unroll
            for (int k = 0; k < ORD; ++k) acc = fma(P[i][k], s[k], acc);
            nx[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < ORD; ++i) s[i] = nx[i];
    }
#pragma unroll
    for (int i = 0; i < ORD; ++i) zi[sess * ORD + i] = s[i];
}

// Parallel carry.  Phi = M^2048 is tiny for this filter (|Phi| = 3.6e-5 although the transient gain of M is 6e5), so the
// state at a chunk start only depends on the last few chunk responses: S_c = e_{c-1} + Phi (e_{c-2} + Phi (e_{c-3} + ...)).
// One thread per (session, chunk) evaluates the recurrence over the last `depth` chunks only, with depth chosen on the
// host so that |Phi|^depth < 2^-70 (what is dropped is far below one ulp of the state); chunks c <= depth start from the
// true initial state, exactly as the sequential carry does.  The sequential version walked 4 688 chunks per session on one
// thread: 1.7 ms, more than both filter passes together.
template <int ORD>
__global__ void k_lp_carry_par(const double* __restrict__ e, double* __restrict__ start, const double* __restrict__ phi,
                               const double* __restrict__ zi_in, double* __restrict__ zi_out, int n_chunks, int n_sessions, int depth) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_sessions * (n_chunks + 1)) return;
    const int sess = (int)(idx / (n_chunks + 1)), ci = (int)(idx - (long long)sess * (n_chunks + 1));
    double P[ORD][ORD], s[ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) {
#pragma unroll
        for (int k = 0; k < ORD; ++k) P[i][k] = __ldg(phi + i * ORD + k);
    }
    const int j0 = ci - depth > 0 ? ci - depth : 0;
#pragma unroll
    for (int i = 0; i < ORD; ++i) s[i] = (j0 == 0) ? zi_in[sess * ORD + i] : 0.0;
    const double* ep = e + (long long)sess * n_chunks * kLpMaxOrd;
    for (int j = j0; j < ci; ++j) {
        double nx[ORD];
#pragma unroll
        for (int i = 0; i < ORD; ++i) {
            double acc = ep[j * kLpMaxOrd + i];
#pragma unroll
            for (int k = 0; k < ORD; ++k) acc = fma(P[i][k], s[k], acc);
            nx[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < ORD; ++i) s[i] = nx[i];
    }
    if (ci < n_chunks) {
        double* sp = start + ((long long)sess * n_chunks + ci) * kLpMaxOrd;
#pragma unroll
        for (int i = 0; i < ORD; ++i) sp[i] = s[i];
    } else {
#pragma unroll
        for (int i = 0; i < ORD; ++i) zi_out[sess * ORD + i] = s[i];
    }
}

// test hook: the kernel's exp(angle()) on arbitrary inputs (tests/test_gpu_decode.py checks it against numpy and mpmath)
__global__ void k_exp_angle(const double* __restrict__ im, const double* __restrict__ re, long long n, double* __restrict__ out) {
    __shared__ double s_tab[kEaTabLen];
    exp_angle_load_table(s_tab);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double vi[1] = {im[i]}, vr[1] = {re[i]};
        double o[1];
        exp_angle_n<1>(s_tab, vi, vr, o);
        out[i] = o[0];
    }
}

int exp_angle_run(const double* im, const double* re, long long n, double* out, cudaStream_t st) {
    if (n <= 0) return SGS_OK;
    k_exp_angle<<<ceil_div(n, 256), 256, 0, st>>>(im, re, n, out);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
int gl_blocks_run(const double* logmel, const double* noise, unsigned long long seed, double* blocks, const GlNodeTables& tab,
                  int n_sessions, int n_frames, int n_mels, int first_frame, int iters, long long ring_base, int ring_len,
                  cudaStream_t st) {
    const long long n_items = (long long)n_sessions * (n_frames - first_frame);
    if (n_items <= 0) return SGS_OK;
    ProfScope ps(kProfGlBlocks, st);
    // one CTA of 16 warps per SM: 128 registers per thread = the whole register file, 226 KB of shared memory (the tables once +
    // 13.4 KB per warp: 32 blocks in flight per SM)
    constexpr int W = 16;
    const size_t smem = sizeof(double) * kFft + sizeof(cplx) * (kG8SLen + 2 * kG8BufCplx) + (sizeof(double2) + sizeof(int2)) * kG8SLen +
                        sizeof(double) * kEaTabLen + sizeof(G8WarpSmem) * W;
    const long long n_pairs = (n_items + 1) / 2, want = (n_pairs + W - 1) / W;
    // grid: 1 / 2 / 4 / 8 / 16 CTAs per SM in all measured 28.13 / 28.04 / 27.96 / 27.89 / 27.75 ms; 8 / 12 / 16 warps per CTA
    // 31.09 / 28.01 / 27.98 ms - the kernel does not speed up past 12 resident warps, so it is not short of warps
    static const int grid_mult = getenv("SGS_GL_GRID_MULT") ? atoi(getenv("SGS_GL_GRID_MULT")) : 16;
    const int grid = (int)(want < 148LL * grid_mult ? want : 148LL * grid_mult);
    const G8Tables t8{tab.window, tab.tw_t, tab.tw_full, tab.inv_idx, tab.inv_w, tab.log_mels};
    static unsigned long long optin = 0;
    SGS_CUDA(smem_optin(k_gl_blocks8<W, 1>, smem, &optin));
    k_gl_blocks8<W, 1><<<grid, W * 32, smem, st>>>(logmel, noise, seed, blocks, t8, n_frames, n_mels, first_frame, iters, n_items,
                                                   ring_base, ring_len);
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

// zi: initial low-pass state per session; zi_out receives the final one (may not alias zi when carry_depth > 0).
// carry_depth > 0 selects the parallel carry (see k_lp_carry_par), 0 the sequential one.
int gl_emit_run(const double* blocks, const int* pos, const double* ola_window, double* v, double* states, double* zi,
                double* zi_out, int carry_depth, const double* phi, const LpCoefs& c, double norm_div, short* pcm, double* filtered,
                int n_sessions, int n_frames, int first_frame, long long n_out, int chunk, int n_chunks, cudaStream_t st) {
    if (chunk != kLpChunk) { set_error("low-pass chunk must be %d samples", kLpChunk); return SGS_ERR_ARG; }
    if (n_out <= 0 || n_frames <= first_frame) return SGS_OK;
    { ProfScope ps(kProfGlOla, st); k_gl_ola<<<dim3(ceil_div(n_frames - first_frame, kOlaFrames), n_sessions), 256, 0, st>>>(blocks, pos, ola_window, v, n_frames, first_frame, n_out); }
    SGS_LAUNCHED();
    ProfScope ps_lp(kProfLowpass, st);
    // states: [2][sessions][chunks][8] - zero-state chunk responses, then true chunk start states
    double* e_states = states;
    double* start_states = states + (size_t)n_sessions * n_chunks * kLpMaxOrd;
    const int n_groups = ceil_div(n_chunks, 32);
    const dim3 grid(ceil_div(n_groups, kLpWarps), n_sessions);
#define SGS_LP(ORD)                                                                                              \
    do {                                                                                                         \
        k_lp_pass<ORD, false><<<grid, kLpWarps * 32, 0, st>>>(v, start_states, e_states, pcm, filtered, c, norm_div, n_out, n_chunks, n_groups); \
        SGS_LAUNCHED();                                                                                          \
        if (carry_depth > 0)                                                                                     \
            k_lp_carry_par<ORD><<<ceil_div((long long)n_sessions * (n_chunks + 1), 128), 128, 0, st>>>(e_states, start_states, phi, zi, zi_out, n_chunks, n_sessions, carry_depth); \
        else {                                                                                                   \
            k_lp_carry<ORD><<<ceil_div(n_sessions, 32), 32, 0, st>>>(e_states, start_states, phi, zi, n_chunks, n_sessions); \
            if (zi_out != zi) cudaMemcpyAsync(zi_out, zi, sizeof(double) * n_sessions * ORD, cudaMemcpyDeviceToDevice, st); \
        }                                                                                                        \
        SGS_LAUNCHED();                                                                                          \
        k_lp_pass<ORD, true><<<grid, kLpWarps * 32, 0, st>>>(v, start_states, e_states, pcm, filtered, c, norm_div, n_out, n_chunks, n_groups); \
        SGS_LAUNCHED();                                                                                          \
    } while (0)
    switch (c.ord) {
        case 1: SGS_LP(1); break; case 2: SGS_LP(2); break; case 3: SGS_LP(3); break; case 4: SGS_LP(4); break;
        case 5: SGS_LP(5); break; case 6: SGS_LP(6); break; case 7: SGS_LP(7); break; default: SGS_LP(8); break;
    }
#undef SGS_LP
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
