// Feature-extraction kernels.  See feat.cuh for the reference semantics being restated.
//
// Data layout in HBM
//   x      [session][t][channel]   fp32 or fp64, time-major exactly as the reference hands it over
//                                  (samples x channels)
//   feat   [session][window][channel]  fp64 log-power, un-stacked (the stacked view is a gather)
//   state  [chunk][state][stream]  fp64, stream = session*C + channel; slot j = cascade state at the
//                                  START of time chunk j (slot 0 = the reference's cold/warm start)
//
// Parallel decomposition (k_iir_stages): a CTA is 4 warps = 4 pipeline stages of the cascade (6 biquads
// each) for 32 streams (lane = stream).  Stage w consumes the 16-sample batch stage w-1 produced in the
// previous iteration through a double-buffered shared-memory hand-off; one bar.sync per iteration.
// Why stages across warps: with warp-uniform roles every coefficient is a uniform-register operand.
// Measured on B200 (tools/pipe_peak.cu): DFMA with three distinct 64-bit register operands issues every
// 3 cycles (12.3 T/s), with a uniform operand every 2 (18.5 T/s); a thread that owns all 24 sections needs
// 78 coefficient doubles and stalls on LDCU re-loads (first version: 37 % of the pipe), and a lane-systolic
// version with per-lane coefficient registers is capped by the register-file limit (second version: 50 %).
// Stage 0 reads the input coalesced (32 consecutive channels of one sample = 128 B per warp load), one
// batch ahead; the last stage keeps a prefix-sum ring of y^2 and closes windows once per batch.
//
// Time is cut into chunks that are made independent by an exact scan over the 48-state recurrence:
//   k_iir_init    slot 0 = state before the first sample (cold start on x[0], offline.py:51-62)
//   pass 1        k_iir_stages<STATE>: zero-state run over the last `horizon` samples of chunk j -> slot j+1
//   k_iir_carry   slot j+1 = Phi(L) slot j + e_j   (only when horizon >= L, i.e. Phi(L) is not negligible)
//   pass 2        k_iir_stages<FEAT>: re-run each chunk from its slot, fused with window energy + log
// `horizon` is chosen on the host from the actual transition matrix so that |A^horizon| is below the
// requested tolerance (default 2^-70, far under one fp64 ulp of the state), see sgs/features.py.
#include "kernels.cuh"

namespace sgs {

template <typename T> __device__ __forceinline__ double ld_d(const T* p) { return (double)__ldg(p); }

constexpr int kThreads = 128;

// ------------------------------------------------------------------------------------------------
// slot 0: the state the reference starts from (local/offline.py:39-62, FrameBuffer.py:87-98):
// filter 0: zi * x[0]; middle filter(s): zi * (first output of the previous filter); last: warm state.
// ------------------------------------------------------------------------------------------------
template <int NB, typename TIn>
__global__ void k_iir_init(const TIn* __restrict__ x, double* __restrict__ state,
                           const __grid_constant__ FeatCoefs cf, const __grid_constant__ FeatGeom g) {
    const int stream = blockIdx.x * blockDim.x + threadIdx.x;
    if (stream >= g.n_streams) return;
    constexpr int NF = NB / kSecPerFilter;
    const int sess = stream / g.n_channels, ch = stream - sess * g.n_channels;
    double v = ld_d(x + (long long)sess * g.session_stride + ch);
    double* out = state + stream;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double z0[kSecPerFilter], z1[kSecPerFilter];
#pragma unroll
        for (int s = 0; s < kSecPerFilter; ++s) {
            const int i = f * kSecPerFilter + s;
            z0[s] = (f == NF - 1) ? cf.zi_warm[s][0] : cf.zi[i][0] * v;
            z1[s] = (f == NF - 1) ? cf.zi_warm[s][1] : cf.zi[i][1] * v;
            out[(long long)(2 * i) * g.state_stride] = z0[s];
            out[(long long)(2 * i + 1) * g.state_stride] = z1[s];
        }
        // first output of this filter (state left untouched: the main kernel processes sample 0 again)
#pragma unroll
        for (int s = 0; s < kSecPerFilter; ++s) v = fma(cf.c[f * kSecPerFilter + s][0], v, z0[s]);
    }
}

enum { kModeState = 0, kModeFeat = 1 };

constexpr int kStages = 4;                      // warps per CTA = pipeline stages of the cascade
constexpr int kStagesC = kStages;
constexpr int kBatch = 16;                      // samples handed from stage to stage per iteration
#ifndef SGS_LOAD_AFTER
#define SGS_LOAD_AFTER 2
#endif
constexpr int kLoadAfter = SGS_LOAD_AFTER;      // stage 0 issues the next batch's loads after this many samples of the current one
constexpr int kStreamsPerBlock = 32;            // lane = stream

// barrier among the 4 stage warps of one pipeline (named barrier `id`, 128 threads; id 0 when the CTA is one pipeline)
__device__ __forceinline__ void pipe_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// One pipeline stage = BPS consecutive biquads of the cascade, run by one warp for 32 streams.
// Coefficients are compile-time offsets into the __grid_constant__ parameter block, so they reach the
// FP64 pipe as uniform-register / constant operands: a DFMA with three distinct 64-bit REGISTER operands
// only issues every 3 cycles on sm_100 (measured, tools/pipe_peak.cu), with a uniform operand every 2.
// Where a segment's cascade state comes from and goes to: element i of this lane's stream sits at base[i * stride]
// (base already includes the lane / stream offset); in == nullptr starts from the zero state.
struct SegState { const double* in; long long in_stride; double* out; long long out_stride; };

// Sections per stage.  The first stage also loads and converts the input, the last one also keeps the prefix ring and closes
// the windows (a log per window): with 6 sections each the middle stages waited 23 % of their time at the barrier for those
// two (barrier stall samples per stage: 1.4 / 5.9 / 5.9 / 0.7 % of the kernel's).
#ifndef SGS_SPLIT_FIRST
#define SGS_SPLIT_FIRST 4      // sections of stage 0 per 24
#endif
#ifndef SGS_SPLIT_LAST
#define SGS_SPLIT_LAST 4       // sections of the last stage per 24
#endif
__host__ __device__ constexpr int stage_sections(int nb, int stage) {
    const int first = nb * SGS_SPLIT_FIRST / 24, last = nb * SGS_SPLIT_LAST / 24, mid = nb - first - last;
    return stage == 0 ? first : stage == kStagesC - 1 ? last : stage == 1 ? (mid + 1) / 2 : mid / 2;
}
__host__ __device__ constexpr int stage_first(int nb, int stage) {
    int f = 0;
    for (int s = 0; s < stage; ++s) f += stage_sections(nb, s);
    return f;
}

template <int NB, bool MONIC, int MODE, int RING, int STAGE, typename TIn>
__device__ __forceinline__ void run_stage(const TIn* __restrict__ x, double* __restrict__ feat,
                                          const SegState ss, const int* __restrict__ starts,
                                          const double* __restrict__ zero_fill_resp, const FeatCoefs& cf,
                                          const FeatGeom& g, double* __restrict__ smem, const int bar_id, const int group,
                                          const bool at_stream_start, const long long t_begin, const int len, const int k_lo,
                                          const int k_hi) {
    constexpr int BPS = stage_sections(NB, STAGE), FIRST = stage_first(NB, STAGE);
    constexpr bool LAST = STAGE == kStages - 1;
    const int lane = threadIdx.x & 31;
    int stream = group * kStreamsPerBlock + lane;
    const bool live = stream < g.n_streams;
    if (!live) stream = g.n_streams - 1;
    const int sess = stream / g.n_channels, ch = stream - sess * g.n_channels;
    const long long C = g.n_channels;

    // stage hand-off buffers: buf[s][parity][kBatch][32] (output of stage s), then the prefix ring (FEAT)
    double* buf_out = smem + (size_t)STAGE * 2 * kBatch * 32;
    const double* buf_in = smem + (size_t)(STAGE > 0 ? STAGE - 1 : 0) * 2 * kBatch * 32;
    double* ring = smem + (size_t)(kStages - 1) * 2 * kBatch * 32;         // [RING][32]

    // state of this stage's sections for this lane's stream
    double z0[BPS], z1[BPS];
    {
#pragma unroll
        for (int p = 0; p < BPS; ++p) {
            const int i = FIRST + p;
            z0[p] = ss.in ? ss.in[(long long)(2 * i) * ss.in_stride] : 0.0;
            z1[p] = ss.in ? ss.in[(long long)(2 * i + 1) * ss.in_stride] : 0.0;
        }
    }

    // stage 0: input prefetch, one batch ahead, coalesced (32 consecutive channels of one sample per warp load)
    const TIn* xrow = x + (long long)sess * g.session_stride + ch + t_begin * C;
    const long long t_last = g.n_samples - 1 - t_begin;                    // clamp: never read past the session
    // The batch being filtered is held as doubles, converted at the END of the iteration that loaded it.  Converting at the
    // point of use put the F2F of sample 0 on the scoreboard the compiler had just given to the next batch's 16 loads, so the
    // first DFMA of every iteration waited for DRAM (long_scoreboard on that one instruction: 6.3 % of the kernel's stall
    // samples, a quarter of stage 0's time - and the other three stages wait for stage 0 at the barrier).
    double xq[STAGE == 0 ? kBatch : 1];
    if (STAGE == 0) {
#pragma unroll
        for (int u = 0; u < kBatch; ++u) xq[u] = ld_d(xrow + (long long)(u < t_last ? u : t_last) * C);
    }

    // last stage: prefix-sum ring + window closing (see k_iir_stages header)
    constexpr int kRebase = 4096;
    double P = 0.0;
    int ke = k_lo, next_end = 0x7fffffff, after_end = 0x7fffffff, rebase_m = -0x40000000, zf_lo = 0;
    double rebase_off = 0.0;
    double* fp = feat + ((long long)sess * g.n_windows) * C + ch;
    if (LAST && MODE == kModeFeat) {
        next_end = (int)((long long)starts[k_lo] + g.window_len - t_begin);
        if (k_lo + 1 < k_hi) after_end = (int)((long long)starts[k_lo + 1] + g.window_len - t_begin);
        if (at_stream_start && g.t_first < 0) {
            // online framing starts inside the warm-start zero fill of the last filter (FrameBuffer.py:95-98): its
            // response there is a session-independent table; seed the ring with its prefix sums
            zf_lo = g.t_first;
            for (int t = g.t_first; t < 0; ++t) {
                const double r = zero_fill_resp[t + g.zero_fill];
                P = fma(r, r, P);
                ring[(t & (RING - 1)) * 32 + lane] = P;
            }
        }
    }

    const int n_batches = (len + kBatch - 1) / kBatch;
    for (int it = 0; it < n_batches + kStages - 1; ++it) {
        const int bidx = it - STAGE;
        if (bidx >= 0 && bidx < n_batches) {
            const int m0 = bidx * kBatch;                                  // first sample (relative) of this batch
            TIn xn[STAGE == 0 ? kBatch : 1];                                // raw samples of the next batch (a float64 recording keeps all its bits)
            const double* in = buf_in + (size_t)((it - 1) & 1) * kBatch * 32 + lane;
            double* out = buf_out + (size_t)(it & 1) * kBatch * 32 + lane;
            // one sample through this stage's sections
            auto step = [&](const int u, double v) {
#pragma unroll
                for (int p = 0; p < BPS; ++p) {
                    const int i = FIRST + p;
                    double y;
                    if (MONIC && (i % kSecPerFilter) != 0) {
                        y = v + z0[p];
                        z0[p] = fma(cf.c[i][1], v, fma(-cf.c[i][3], y, z1[p]));
                        z1[p] = fma(-cf.c[i][4], y, v);
                    } else {
                        y = fma(cf.c[i][0], v, z0[p]);
                        z0[p] = fma(cf.c[i][1], v, fma(-cf.c[i][3], y, z1[p]));
                        z1[p] = fma(-cf.c[i][4], y, cf.c[i][2] * v);
                    }
                    v = y;
                }
                if (!LAST) out[u * 32] = v;
                else if (MODE == kModeFeat) {
                    P = fma(v, v, P);
                    ring[((m0 + u) & (RING - 1)) * 32 + lane] = P;
                }
            };
            if (STAGE == 0) {
                // Stage 0 runs its (few) sections over the whole batch in one straight line: its coefficients then stay in
                // uniform registers for the iteration, and the next batch's 16 loads are issued AFTER the first samples.  With
                // the loads in front, the first DFMA of each half-batch waited for DRAM: the scoreboard of the uniform
                // coefficient load it needs was shared with the loads just issued (6.6 % of the kernel's stall samples on
                // that one instruction, a quarter of this stage's time).
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    if (u == kLoadAfter) {
#pragma unroll
                        for (int w = 0; w < kBatch; ++w) {
                            const long long m = (long long)m0 + kBatch + w;
                            xn[w] = __ldg(xrow + (m < t_last ? m : t_last) * C);
                        }
                    }
                    step(u, xq[u]);
                }
            } else {
                // two half-batches of 8 samples: keeps the stage's loop body inside the per-scheduler instruction cache
                // while the 4 warps of a pipeline execute 4 different bodies
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int uu = 0; uu < kBatch / 2; ++uu) {
                        const int u = h * (kBatch / 2) + uu;
                        step(u, in[u * 32]);
                    }
                }
            }
            if (STAGE == 0) {
#pragma unroll
                for (int u = 0; u < kBatch; ++u) xq[u] = (double)xn[u];
            }
            if (LAST && MODE == kModeFeat) {
                const int done = m0 + kBatch;                              // samples [0, done) have their prefix in the ring
                while (next_end <= done) {
                    const int e = next_end, s0 = e - g.window_len;
                    double hi = ring[((e - 1) & (RING - 1)) * 32 + lane];
                    const double lo = (s0 - 1 < zf_lo) ? 0.0 : ring[((s0 - 1) & (RING - 1)) * 32 + lane];
                    if (s0 - 1 < rebase_m && e - 1 >= rebase_m) hi += rebase_off;      // window straddles the last re-base
                    if (live) fp[(long long)ke * C] = log((hi - lo) + 0.01);
                    ++ke;
                    next_end = after_end;
                    after_end = (ke + 1 < k_hi) ? (int)((long long)starts[ke + 1] + g.window_len - t_begin) : 0x7fffffff;
                }
                if (done - rebase_m >= kRebase) {
                    // P restarts from 0 so that hi - lo never cancels more than ~kRebase/window_len ulps
                    rebase_m = done;
                    rebase_off = P;
                    P = 0.0;
                }
            }
        }
        pipe_sync(bar_id);
    }

    if (MODE == kModeState && live) {
#pragma unroll
        for (int p = 0; p < BPS; ++p) {
            const int i = FIRST + p;
            ss.out[(long long)(2 * i) * ss.out_stride] = z0[p];
            ss.out[(long long)(2 * i + 1) * ss.out_stride] = z1[p];
        }
    }
}

#define SGS_RUN_STAGE(S) run_stage<NB, MONIC, MODE, RING, S, TIn>(x, feat, ss, starts, zero_fill_resp, cf, g, smem_p, bar_id, group, at_start, t_begin, len, k_lo, k_hi)
#define SGS_DISPATCH_STAGE(stage_warp, rot)                                                        \
    switch (((stage_warp) + (rot)) & (kStages - 1)) {                                              \
        case 0: SGS_RUN_STAGE(0); break;                                                           \
        case 1: SGS_RUN_STAGE(1); break;                                                           \
        case 2: SGS_RUN_STAGE(2); break;                                                           \
        default: SGS_RUN_STAGE(3); break;                                                          \
    }

// grid = (ceil(streams/32), chunks); block = 4 warps = 4 pipeline stages.
// FEAT mode: the last stage keeps a running prefix sum P of y^2 per stream and writes it to a ring in shared
// memory (slot = sample index & (RING-1)); a window [s, e) is P[e-1] - P[s-1], closed once per batch, all 32
// streams of the warp at once (coalesced 256 B store of the log-power row segment).
template <int NB, bool MONIC, int MODE, int RING, typename TIn>
__global__ void __launch_bounds__(kStages * 32)
k_iir_stages(const TIn* __restrict__ x, double* __restrict__ feat, double* __restrict__ state,
             const long long* __restrict__ bounds, const int* __restrict__ kfirst, const int* __restrict__ starts,
             const double* __restrict__ zero_fill_resp, const __grid_constant__ FeatCoefs cf,
             const __grid_constant__ FeatGeom g) {
    extern __shared__ double smem[];
    const int j = blockIdx.y, group = blockIdx.x;
    long long t_begin = bounds[j], t_end;
    int k_lo = 0, k_hi = 0;
    bool truncated = false;
    if (MODE == kModeState) {
        t_end = bounds[j + 1];
        if (t_end - g.horizon > t_begin) { t_begin = t_end - g.horizon; truncated = true; }   // truncated zero-state pass
    } else {
        k_lo = kfirst[j]; k_hi = kfirst[j + 1];
        if (k_lo >= k_hi) return;
        t_end = (long long)starts[k_hi - 1] + g.window_len;                // run on until the last owned window closes
    }
    const int len = (int)(t_end - t_begin);
    // in STATE mode the host guarantees len % kBatch == 0 so that the end state is taken exactly at t_end
    const int lane = threadIdx.x & 31;
    int stream = group * kStreamsPerBlock + lane;
    if (stream >= g.n_streams) stream = g.n_streams - 1;
    double* slot_j = state + ((long long)j * (2 * NB)) * g.state_stride + stream;
    SegState ss;
    ss.in = (MODE == kModeState && (j > 0 || truncated)) ? nullptr : slot_j;
    ss.in_stride = g.state_stride;
    ss.out = slot_j + (long long)(2 * NB) * g.state_stride;
    ss.out_stride = g.state_stride;
    const bool at_start = j == 0;
    double* smem_p = smem;
    const int bar_id = 0;
    // rotate the stage -> warp (= scheduler) assignment with the block index so that co-resident CTAs do not
    // stack all their heaviest stages on the same scheduler
    SGS_DISPATCH_STAGE(threadIdx.x >> 5, blockIdx.x + blockIdx.y)
}

// ---- balanced pieces -------------------------------------------------------------------------------------------------------
// The time lines of all stream groups are laid end to end and cut into P equal pieces, P a multiple of the SM count, one CTA
// per piece: with (groups x chunks) CTAs the 148 SMs held 2 or 3 CTAs each and the kernel ran at the pace of the 3-CTA SMs
// (SMs active 88 % of the time, profiles/ncu_iir_feat32_r01.txt; 37 sessions - 444 CTAs - took exactly as long as 32).  A piece
// is one or two segments (the end of one group's recording and the start of the next one's).  A segment that does not start
// at t = 0 gets its start state from the zero-state pass over the `horizon` samples before it (or, nearer than that to the
// beginning, from the true initial state), written to its own slot seg_state[segment][2 NB][32].
// PIPES pipelines per CTA: pipeline p = warps {p, p + PIPES, p + 2 PIPES, p + 3 PIPES}.  With PIPES = 4 the four stage warps
// of a pipeline have warp indices congruent mod 4, i.e. they share one of the SM's four schedulers: a stage waiting at the
// pipeline's barrier hands its issue slots to exactly the stage it is waiting for, instead of idling a scheduler while
// another one is overloaded (the per-batch barrier was the largest stall with one stage per scheduler).
template <int NB, bool MONIC, int MODE, int RING, int PIPES, typename TIn>
__global__ void __launch_bounds__(PIPES * kStages * 32, 1)
k_iir_pieces(const TIn* __restrict__ x, double* __restrict__ feat, const double* __restrict__ init_state /*[2NB][streams]*/,
             double* __restrict__ seg_state /*[segment][2NB][32]*/, const FeatSeg* __restrict__ segs, const int* __restrict__ piece_first,
             int n_pieces, const int* __restrict__ starts, const double* __restrict__ zero_fill_resp, const __grid_constant__ FeatCoefs cf,
             const __grid_constant__ FeatGeom g) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pipe = warp % PIPES, stage_warp = warp / PIPES;
    const int piece = blockIdx.x * PIPES + pipe;
    if (piece >= n_pieces) return;
    constexpr int pipe_doubles = (kStages - 1) * 2 * kBatch * 32 + (MODE == kModeFeat ? RING * 32 : 0);
    double* smem_p = smem + (size_t)pipe * pipe_doubles;
    const int bar_id = PIPES > 1 ? 1 + pipe : 0;
    for (int si = piece_first[piece]; si < piece_first[piece + 1]; ++si) {
        const FeatSeg sg = segs[si];
        const int group = sg.group;
        int stream = group * kStreamsPerBlock + lane;
        if (stream >= g.n_streams) stream = g.n_streams - 1;
        double* slot = seg_state + ((long long)si * (2 * NB)) * 32 + lane;
        const double* init = init_state + stream;
        SegState ss;
        long long t_begin;
        int len, k_lo = sg.k_lo, k_hi = sg.k_hi;
        bool at_start;
        if (MODE == kModeState) {
            if (sg.t_begin == 0) continue;                                   // starts from the initial state: nothing to compute
            t_begin = sg.warm_begin;
            len = (int)(sg.t_begin - sg.warm_begin);
            // from the true initial state, from the modal tail sum of the far past (k_iir_tail, same slot), or from zero
            ss.in = sg.tail ? slot : sg.warm_begin == 0 ? init : nullptr; ss.in_stride = sg.tail ? 32 : g.state_stride;
            ss.out = slot; ss.out_stride = 32;
            at_start = false;
        } else {
            if (k_lo >= k_hi) continue;
            t_begin = sg.t_begin;
            len = (int)((long long)starts[k_hi - 1] + g.window_len - t_begin);
            ss.in = sg.t_begin == 0 ? init : slot; ss.in_stride = sg.t_begin == 0 ? g.state_stride : 32;
            ss.out = nullptr; ss.out_stride = 0;
            at_start = sg.t_begin == 0;
        }
        SGS_DISPATCH_STAGE(stage_warp, piece)
        pipe_sync(bar_id);                                                   // the hand-off buffers and the ring are reused
    }
}
#undef SGS_DISPATCH_STAGE
#undef SGS_RUN_STAGE

// ---- modal tail (sgs/modal.py) -----------------------------------------------------------------------------------------------
// Start state of a segment's warm-up at t_near = t_begin - near_len from the samples before it: for each of the few modes that
// outlive near_len, c_m = sum_k lambda_m^k x[t_near - 1 - k] over the mode's own horizon, then state = M (Re c, Im c).  The
// complex one-pole sum runs as its real second-order form (a damped Goertzel recurrence): y(n) = 2 Re(lambda) y(n-1) -
// |lambda|^2 y(n-2) + x(n), c = y(n) - conj(lambda) y(n-1) - two FMAs per sample and mode against the cascade's ~99, and only
// two constants per mode, so every DFMA has a uniform-register operand (a table of 32 powers per mode was tried first: 256
// distinct constants per block overflow the uniform register file and came back as R2UR moves, one per DFMA).
// One CTA per segment, lane = stream.  Modes come in groups of 4 with a common length; the (group, block of 32 samples) pairs of
// all groups, laid end to end, are dealt to the 16 warps in equal runs (cut on the host); a warp runs its blocks oldest first
// from y = 0, moves its sum to t_near (lambda^(32 first block)), the sums are added in warp order (bit-reproducible) and
// multiplied by the real 2 nb x 2 n_modes matrix M.
template <int G, typename TIn>
__device__ __forceinline__ void tail_group(const TIn* __restrict__ xs, const long long C, const long long t_near, const int d_lo,
                                           const int d_hi, const int warp, const int lane, double* __restrict__ part, const TailTab& tt) {
    double y1[4] = {0.0, 0.0, 0.0, 0.0}, y2[4] = {0.0, 0.0, 0.0, 0.0};     // y(n-1), y(n-2)
    if (d_hi > d_lo) {
        const TIn* xp = xs + (t_near - (long long)kTailBlock * d_hi) * C;  // blocks d_hi-1 .. d_lo are contiguous in time
        TIn xv[kTailBlock];
#pragma unroll
        for (int j = 0; j < kTailBlock; ++j) { xv[j] = __ldg(xp); xp += C; }
#pragma unroll 1
        for (int d = d_hi - 1; d >= d_lo; --d) {
            double xd[kTailBlock];
#pragma unroll
            for (int j = 0; j < kTailBlock; ++j) xd[j] = (double)xv[j];
            if (d > d_lo) {                                                // the next block's samples, requested before this one's arithmetic
#pragma unroll
                for (int j = 0; j < kTailBlock; ++j) { xv[j] = __ldg(xp); xp += C; }
            }
#pragma unroll
            for (int j = 0; j < kTailBlock; ++j) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double y = fma(tt.rec[4 * G + q][0], y1[q], fma(tt.rec[4 * G + q][1], y2[q], xd[j]));
                    y2[q] = y1[q];
                    y1[q] = y;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int m = 4 * G + q;
        const double cr = fma(-tt.lam[m][0], y2[q], y1[q]), ci = tt.lam[m][1] * y2[q];    // c = y(n) - conj(lambda) y(n-1)
        const double sr = tt.shift[warp][m][0], si = tt.shift[warp][m][1];
        part[(warp * 2 * kTailMaxModes + m) * 32 + lane] = fma(cr, sr, -ci * si);
        part[(warp * 2 * kTailMaxModes + kTailMaxModes + m) * 32 + lane] = fma(cr, si, ci * sr);
    }
}

template <int NS, typename TIn>
__global__ void __launch_bounds__(kTailWarps * 32, 1)
k_iir_tail(const TIn* __restrict__ x, double* __restrict__ seg_state /*[segment][NS][32]*/, const FeatSeg* __restrict__ segs,
           const double* __restrict__ matrix /*[NS][2 n_modes]*/, const double* __restrict__ kappa /*[n_modes][2][NS]*/,
           const double* __restrict__ init_state /*[NS][streams]*/, const __grid_constant__ TailTab tt,
           const __grid_constant__ FeatGeom g) {
    extern __shared__ double smem[];
    double* part = smem;                                                   // [warp][2 kTailMaxModes][32]
    double* tot = smem + kTailWarps * 2 * kTailMaxModes * 32;              // [2 kTailMaxModes][32]
    const FeatSeg sg = segs[blockIdx.x];
    if (!sg.tail) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int stream = sg.group * kStreamsPerBlock + lane;
    if (stream >= g.n_streams) stream = g.n_streams - 1;
    const int sess = stream / g.n_channels, ch = stream - sess * g.n_channels;
    const long long C = g.n_channels;
    const long long t_near = sg.t_begin - tt.near_len;
    const TIn* xs = x + (long long)sess * g.session_stride + ch;
    // a group this warp has no blocks of contributes exact zeros (lo >= hi).  Nearer to the start of the recording than the
    // horizon (tail == 2) the blocks before t = 0 do not exist - the state there is the reference's initial state, added below
    const int have = (int)min(t_near / kTailBlock, (long long)0x7fffffff);
    tail_group<0, TIn>(xs, C, t_near, tt.blk_lo[warp][0], min(tt.blk_hi[warp][0], have), warp, lane, part, tt);
    tail_group<1, TIn>(xs, C, t_near, tt.blk_lo[warp][1], min(tt.blk_hi[warp][1], have), warp, lane, part, tt);
    tail_group<2, TIn>(xs, C, t_near, tt.blk_lo[warp][2], min(tt.blk_hi[warp][2], have), warp, lane, part, tt);
    tail_group<3, TIn>(xs, C, t_near, tt.blk_lo[warp][3], min(tt.blk_hi[warp][3], have), warp, lane, part, tt);
    __syncthreads();
#pragma unroll 1
    for (int v = warp; v < 2 * kTailMaxModes; v += kTailWarps) {
        double a = 0.0;
#pragma unroll 4
        for (int w = 0; w < kTailWarps; ++w) a += part[(w * 2 * kTailMaxModes + v) * 32 + lane];
        tot[v * 32 + lane] = a;
    }
    __syncthreads();
    const int nm = tt.n_modes;
    if (sg.tail == 2) {
        // c_m += lambda_m^t_near (kappa_m . s_init): what the initial state (k_iir_init) still holds in mode m; warp = mode
        for (int m = warp; m < nm; m += kTailWarps) {
            double zr = 1.0, zi = 0.0, br = tt.lam[m][0], bi = tt.lam[m][1];
            for (long long n = t_near; n; n >>= 1) {                      // lambda^t_near by repeated squaring
                if (n & 1) { const double t = fma(zr, br, -zi * bi); zi = fma(zr, bi, zi * br); zr = t; }
                const double t = fma(br, br, -bi * bi); bi = 2.0 * br * bi; br = t;
            }
            const double* kr = kappa + (long long)m * 2 * NS;
            double pr = 0.0, pi = 0.0;
#pragma unroll 1
            for (int i = 0; i < NS; ++i) {
                const double s0 = init_state[(long long)i * g.state_stride + stream];
                pr = fma(__ldg(kr + i), s0, pr);
                pi = fma(__ldg(kr + NS + i), s0, pi);
            }
            tot[m * 32 + lane] += fma(zr, pr, -zi * pi);
            tot[(kTailMaxModes + m) * 32 + lane] += fma(zr, pi, zi * pr);
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int i = warp; i < NS; i += kTailWarps) {
        const double* row = matrix + (long long)i * 2 * nm;
        double a = 0.0;
#pragma unroll 1
        for (int m = 0; m < nm; ++m) a = fma(__ldg(row + m), tot[m * 32 + lane], a);
#pragma unroll 1
        for (int m = 0; m < nm; ++m) a = fma(__ldg(row + nm + m), tot[(kTailMaxModes + m) * 32 + lane], a);
        seg_state[((long long)blockIdx.x * NS + i) * 32 + lane] = a;
    }
}

// ------------------------------------------------------------------------------------------------
// carry (exact mode): slot j+1 = Phi slot j + e_j for j >= 1, in place; slot 1 is already the true state
// (chunk 0 ran from slot 0).  One thread per stream; Phi (NS x NS, row-major) is read warp-uniformly.
// ------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kThreads)
k_iir_carry(double* __restrict__ slots, const double* __restrict__ phi, int n_chunks, int n_streams) {
    const int stream = blockIdx.x * kThreads + threadIdx.x;
    if (stream >= n_streams) return;
    double s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = slots[((long long)NS + i) * n_streams + stream];      // slot 1
    for (int j = 2; j < n_chunks; ++j) {
        double* e = slots + (long long)j * NS * n_streams + stream;
        double r[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) r[i] = e[(long long)i * n_streams];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double acc = r[i];
#pragma unroll
            for (int k = 0; k < NS; ++k) acc = fma(__ldg(phi + i * NS + k), s[k], acc);
            r[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            s[i] = r[i];
            e[(long long)i * n_streams] = r[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// temporal context stacking: out[row][c*(order+1)+tap] = feat[row + first_row - (order-tap)*step][c]
// with exact zeros before the stream start (online) - offline simply starts at first_row = order*step.
// ------------------------------------------------------------------------------------------------
__global__ void k_stack(const double* __restrict__ feat, double* __restrict__ out, int n_windows, int n_channels,
                        int n_rows, int first_row, int order, int step, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int width = n_channels * (order + 1);
    const long long per_sess = (long long)n_rows * width;
    const int sess = (int)(idx / per_sess);
    const long long r = idx - (long long)sess * per_sess;
    const int row = (int)(r / width), col = (int)(r - (long long)row * width);
    const int c = col / (order + 1), tap = col - c * (order + 1);
    const int w = row + first_row - (order - tap) * step;
    out[idx] = (w >= 0) ? feat[((long long)sess * n_windows + w) * n_channels + c] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
template <int NB, bool MONIC, int RING, typename TIn>
static void run_all(const TIn* x, double* feat, double* slots, const double* phi, bool apply_phi, const long long* bounds,
                    const int* kfirst, const int* starts, const double* zf, const FeatCoefs& cf, const FeatGeom& g,
                    cudaStream_t st) {
    { ProfScope ps(kProfIirInit, st); k_iir_init<NB, TIn><<<ceil_div(g.n_streams, 128), 128, 0, st>>>(x, slots, cf, g); }
    SGS_LAUNCHED();
    const int bx = ceil_div(g.n_streams, kStreamsPerBlock);
    constexpr int hand_bytes = (kStages - 1) * 2 * kBatch * 32 * (int)sizeof(double);
    constexpr int feat_bytes = hand_bytes + RING * 32 * (int)sizeof(double);
    static unsigned long long optin = 0;
    if (smem_optin(k_iir_stages<NB, MONIC, kModeFeat, RING, TIn>, feat_bytes, &optin) != cudaSuccess) return;   // surfaces as the launch error
    if (g.n_chunks > 1) {
        {
            ProfScope ps(kProfIirState, st);
            k_iir_stages<NB, MONIC, kModeState, RING, TIn><<<dim3(bx, g.n_chunks - 1), kStages * 32, hand_bytes, st>>>(
                x, feat, slots, bounds, kfirst, starts, zf, cf, g);
        }
        SGS_LAUNCHED();
        if (apply_phi && g.n_chunks > 2) {
            { ProfScope ps(kProfIirCarry, st); k_iir_carry<2 * NB><<<ceil_div(g.n_streams, kThreads), kThreads, 0, st>>>(slots, phi, g.n_chunks, g.n_streams); }
            SGS_LAUNCHED();
        }
    }
    {
        ProfScope ps(kProfIirFeat, st);
        k_iir_stages<NB, MONIC, kModeFeat, RING, TIn><<<dim3(bx, g.n_chunks), kStages * 32, feat_bytes, st>>>(
            x, feat, slots, bounds, kfirst, starts, zf, cf, g);
    }
    SGS_LAUNCHED();
}

constexpr int kPipes = 4;               // pipelines per CTA (see k_iir_pieces)

template <int NB, bool MONIC, int RING, typename TIn>
static void run_pieces(const TIn* x, double* feat, double* init_state, double* seg_state, const FeatSeg* segs, const int* piece_first,
                       int n_pieces, int n_segs, const TailTab* tail, const double* tail_matrix, const int* starts, const double* zf,
                       const FeatCoefs& cf, const FeatGeom& g, cudaStream_t st) {
    { ProfScope ps(kProfIirInit, st); k_iir_init<NB, TIn><<<ceil_div(g.n_streams, 128), 128, 0, st>>>(x, init_state, cf, g); }
    SGS_LAUNCHED();
    constexpr int hand_bytes = kPipes * (kStages - 1) * 2 * kBatch * 32 * (int)sizeof(double);
    constexpr int feat_bytes = hand_bytes + kPipes * RING * 32 * (int)sizeof(double);
    static unsigned long long optin_f = 0, optin_s = 0;
    if (smem_optin(k_iir_pieces<NB, MONIC, kModeFeat, RING, kPipes, TIn>, feat_bytes, &optin_f) != cudaSuccess ||
        smem_optin(k_iir_pieces<NB, MONIC, kModeState, RING, kPipes, TIn>, hand_bytes, &optin_s) != cudaSuccess) return;   // surfaces as the launch error
    const int grid = ceil_div(n_pieces, kPipes);
    if (tail) {
        ProfScope ps(kProfPiecesTail, st);
        constexpr int tail_bytes = (kTailWarps + 1) * 2 * kTailMaxModes * 32 * (int)sizeof(double);
        static unsigned long long optin_t = 0;
        if (smem_optin(k_iir_tail<2 * NB, TIn>, tail_bytes, &optin_t) != cudaSuccess) return;
        k_iir_tail<2 * NB, TIn><<<n_segs, kTailWarps * 32, tail_bytes, st>>>(
            x, seg_state, segs, tail_matrix, tail_matrix + (size_t)2 * NB * 2 * tail->n_modes, init_state, *tail, g);
    }
    if (tail) SGS_LAUNCHED();
    {
        ProfScope ps(kProfPiecesState, st);
        k_iir_pieces<NB, MONIC, kModeState, RING, kPipes, TIn><<<grid, kPipes * kStages * 32, hand_bytes, st>>>(
            x, feat, init_state, seg_state, segs, piece_first, n_pieces, starts, zf, cf, g);
    }
    SGS_LAUNCHED();
    {
        ProfScope ps(kProfPiecesFeat, st);
        k_iir_pieces<NB, MONIC, kModeFeat, RING, kPipes, TIn><<<grid, kPipes * kStages * 32, feat_bytes, st>>>(
            x, feat, init_state, seg_state, segs, piece_first, n_pieces, starts, zf, cf, g);
    }
    SGS_LAUNCHED();
}

template <typename TIn>
static int run_pieces_typed(int n_biquads, bool monic, const TIn* x, double* feat, double* init_state, double* seg_state,
                            const FeatSeg* segs, const int* piece_first, int n_pieces, int n_segs, const TailTab* tail,
                            const double* tail_matrix, const int* starts, const double* zf,
                            const FeatCoefs& cf, const FeatGeom& g, cudaStream_t st) {
#define SGS_RUN(NB, M, RING) run_pieces<NB, M, RING, TIn>(x, feat, init_state, seg_state, segs, piece_first, n_pieces, n_segs, tail, tail_matrix, starts, zf, cf, g, st)
#define SGS_PICK(NB, M)                                             \
    do {                                                            \
        if (g.window_len + kBatch + 1 <= 128) SGS_RUN(NB, M, 128);  \
        else SGS_RUN(NB, M, 256);                                   \
    } while (0)
    if (g.window_len + kBatch + 1 > 256) { set_error("window_len %d too long (max %d)", g.window_len, 255 - kBatch); return SGS_ERR_UNSUPPORTED; }
    if (n_biquads == 24 && monic) SGS_PICK(24, true);
    else if (n_biquads == 24) SGS_PICK(24, false);
    else if (n_biquads == 16 && monic) SGS_PICK(16, true);
    else if (n_biquads == 16) SGS_PICK(16, false);
    else { set_error("unsupported biquad count %d (16 or 24)", n_biquads); return SGS_ERR_UNSUPPORTED; }
#undef SGS_PICK
#undef SGS_RUN
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

int feat_run_pieces(int n_biquads, bool monic, const void* x, bool x_is_f64, double* feat, double* init_state, double* seg_state,
                    const FeatSeg* segs, const int* piece_first, int n_pieces, int n_segs, const TailTab* tail, const double* tail_matrix,
                    const int* starts, const double* zf, const FeatCoefs& cf, const FeatGeom& g, cudaStream_t st) {
    if (x_is_f64)
        return run_pieces_typed<double>(n_biquads, monic, (const double*)x, feat, init_state, seg_state, segs, piece_first, n_pieces,
                                        n_segs, tail, tail_matrix, starts, zf, cf, g, st);
    return run_pieces_typed<float>(n_biquads, monic, (const float*)x, feat, init_state, seg_state, segs, piece_first, n_pieces,
                                   n_segs, tail, tail_matrix, starts, zf, cf, g, st);
}

template <typename TIn>
static int run_typed(int n_biquads, bool monic, const TIn* x, double* feat, double* slots, const double* phi,
                     bool apply_phi, const long long* bounds, const int* kfirst, const int* starts, const double* zf,
                     const double* coef, const FeatCoefs& cf, const FeatGeom& g, cudaStream_t st) {
    (void)coef;
#define SGS_RUN(NB, M, RING) run_all<NB, M, RING, TIn>(x, feat, slots, phi, apply_phi, bounds, kfirst, starts, zf, cf, g, st)
#define SGS_PICK(NB, M)                                             \
    do {                                                            \
        if (g.window_len + kBatch + 1 <= 128) SGS_RUN(NB, M, 128);  \
        else SGS_RUN(NB, M, 256);                                   \
    } while (0)
    if (g.window_len + kBatch + 1 > 256) { set_error("window_len %d too long (max %d)", g.window_len, 255 - kBatch); return SGS_ERR_UNSUPPORTED; }
    if (n_biquads == 24 && monic) SGS_PICK(24, true);
    else if (n_biquads == 24) SGS_PICK(24, false);
    else if (n_biquads == 16 && monic) SGS_PICK(16, true);
    else if (n_biquads == 16) SGS_PICK(16, false);
    else { set_error("unsupported biquad count %d (16 or 24)", n_biquads); return SGS_ERR_UNSUPPORTED; }
#undef SGS_PICK
#undef SGS_RUN
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

int feat_run(int n_biquads, bool monic, const void* x, bool x_is_f64, double* feat, double* slots, const double* phi,
             bool apply_phi, const long long* bounds, const int* kfirst, const int* starts, const double* zf,
             const double* coef, const FeatCoefs& cf, const FeatGeom& g, cudaStream_t st) {
    if (x_is_f64)
        return run_typed<double>(n_biquads, monic, (const double*)x, feat, slots, phi, apply_phi, bounds, kfirst, starts,
                                 zf, coef, cf, g, st);
    return run_typed<float>(n_biquads, monic, (const float*)x, feat, slots, phi, apply_phi, bounds, kfirst, starts, zf,
                            coef, cf, g, st);
}

int stack_run(const double* feat, double* out, int n_sessions, int n_windows, int n_channels, int n_rows, int first_row,
              int order, int step, cudaStream_t st) {
    const long long total = (long long)n_sessions * n_rows * n_channels * (order + 1);
    if (total == 0) return SGS_OK;
    { ProfScope ps(kProfStack, st); k_stack<<<ceil_div(total, 256), 256, 0, st>>>(feat, out, n_windows, n_channels, n_rows, first_row, order, step, total); }
    SGS_LAUNCHED();
    SGS_CUDA(cudaGetLastError());
    return SGS_OK;
}

}  // namespace sgs
