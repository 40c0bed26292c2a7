// High-gamma feature extraction: cascaded Butterworth biquads (DF2T, FP64) as a chunked scan over time,
// fused with the 50 ms window energy, log and (separately) the temporal context stacking.
//
// Reference semantics restated (paths under the reference tree):
//   filters + states ........ local/offline.py:31-97, livenodes/FrameBuffer.py:86-98,139-143
//   recurrence .............. scipy.signal.sosfilt: y=b0*x+z0; z0=b1*x-a1*y+z1; z1=b2*x-a2*y (sections inner loop)
//   window energy + log ..... local/offline.py:99-109, livenodes/ECogFeatCalc.py:118-124
//   stacking ................ local/offline.py:111-116, livenodes/ECogFeatCalc.py:137-144
#pragma once
#include "common.cuh"

namespace sgs {

constexpr int kMaxBiquads = 24;
constexpr int kSecPerFilter = 8;
constexpr int kFifo = 8;             // max simultaneously open windows per stream

// Passed BY VALUE as a __grid_constant__ kernel parameter: the coefficients then sit in the constant
// bank and feed DFMA directly as c[0x0][..] operands (no loads, no registers).
struct FeatCoefs {
    double c[kMaxBiquads][5];        // b0 b1 b2 a1 a2
    double zi[kMaxBiquads][2];       // unit steady-state zi (cold start: scaled by the first sample)
    double zi_warm[kSecPerFilter][2];// last filter after its warm-start zero fill
};

struct FeatGeom {
    long long n_samples;             // T per session
    int n_channels;                  // C
    int n_sessions;
    long long session_stride;        // elements between sessions in x
    int n_streams;                   // C * n_sessions
    int n_windows;                   // windows per session
    int window_len;
    int t_first;                     // first time index chunk 0 walks: -zero_fill (online) or 0 (offline)
    int zero_fill;                   // length of the zero-fill response table
    int n_chunks;
    int horizon;                     // zero-state pass length W (samples before each chunk end)
    int state_stride;                // = n_streams (carry layout [chunk][state][stream])
};

// One segment of a balanced work piece (k_iir_pieces, feat.cu).
struct FeatSeg {
    int group;                  // stream group (32 streams)
    int k_lo, k_hi;             // windows owned: those that START inside the segment
    int tail;                   // 1: the warm-up starts from the modal tail sum k_iir_tail left in the segment's state slot
    long long t_begin;          // first sample of the segment
    long long warm_begin;       // STATE pass: first sample of the warm-up run (0 = from the true initial state)
};

// Modal tail of the cascade (sgs/modal.py): the few pole pairs that outlive `near_len` samples, summed directly over the far
// past of a segment start.  Passed BY VALUE (__grid_constant__): the recurrence constants reach the FP64 pipe as uniform
// operands.  Mode m is live for mode_blocks[m] blocks of kTailBlock samples counted back from t_near = t_begin - near_len;
// modes are sorted by that length (longest first) and come in groups of 4 of equal length.
constexpr int kTailBlock = 32, kTailMaxModes = 16, kTailWarps = 16;
struct TailTab {
    double rec[kTailMaxModes][2];                    // 2 Re(lambda), -|lambda|^2: the real second-order form of the one-pole sum
    double lam[kTailMaxModes][2];                    // lambda (re, im)
    double shift[kTailWarps][kTailMaxModes][2];      // lambda^(32 blk_lo[w][m / 4]): moves warp w's sum to t_near
    int blk_lo[kTailWarps][kTailMaxModes / 4];       // warp w sums blocks [blk_lo, blk_hi) of mode group g; block 0 ends at t_near
    int blk_hi[kTailWarps][kTailMaxModes / 4];
    int mode_blocks[kTailMaxModes];
    int n_modes, near_len, far_len;
};

}  // namespace sgs
