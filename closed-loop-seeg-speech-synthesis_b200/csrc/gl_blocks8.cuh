// k_gl_blocks8: the node-semantics Griffin-Lim block synthesis (see gl_node.cu for the reference semantics) with the
// FFTs held in REGISTERS.
//
// Why: the first version kept every FFT stage in shared memory (4 read+write passes per 128-point transform plus
// twiddle loads) and was bound by shared-memory wavefronts and bank conflicts - its run time did not move between 16,
// 20 and 24 resident warps per SM (profiles/ncu_gl_blocks_*_r01.txt).  Here one 128-point complex transform lives in
// the registers of 8 lanes, 16 points each:
//     n = l + 8m  (l = lane in group, m < 16)        16-point FFT over m, in registers        -> Y[l][k1]
//     Y[l][k1] *= W128^(l k1)                        one table load per point
//     transposition through shared memory            the ONLY exchange: 16 stores + 16 loads of 16 B per lane
//     k = k1 + 16 k2                                 8-point FFT over l, in registers, for the lane's two rows k1
// The rows are dealt so that the real-FFT partner bins k and 128-k sit in the SAME lane: lane l >= 1 owns rows l and
// 16-l (partner of (l, k2) is (16-l, 7-k2)); lane 0 owns the two self-paired rows 0 and 8.  The phase step
// Z = S * exp(angle(X)) (real, quirk Q1) and the inverse split are therefore lane-local; lane 0's different pairing is
// expressed with register selects so that the expensive exp(angle) code runs without divergence.  The inverse runs
// the mirrored pipeline (8-point FFT over k2, twiddle, transposition, 16-point FFT over k1) on conjugated data.
// A warp works on two blocks at once: group g = lane / 8 handles block g / 2, STFT frame g % 2 (offsets 0 and 160).
//
// The WAVEFORM STAYS IN REGISTERS across iterations.  The first register-FFT version wrote the synthesis result to shared
// memory and read it back for the next analysis; it was co-limited by the shared-memory data pipe (64 % busy: 832 wavefronts
// per warp-iteration - every LDS.128 / STS.128 of a warp is 4 wavefronts, broadcast across the four 8-lane groups does not
// merge) and the FP64 pipe (54 %).  A lane's synthesis output r[l8 + 8m] is exactly the sample pair its next analysis reads,
// except in the overlap of the two STFT frames (block samples 160..255): frame 0 holds them at m = 10..15, frame 1 at
// m = 0..5, SAME l8.  Frame 1 therefore keeps point m in register slot (m + 10) mod 16 - a circular shift of the 16-point
// transform's input, i.e. a phase W16^(6 k1) on its output, folded into frame 1's own twiddle table W128^((l + 48) k1) - so
// that both frames hold the overlap in slots 10..15 and ONE xor-8 shuffle + add of 6 complex registers replaces the store /
// load-add-store / load round trip; the synthesis and analysis windows (the same two doubles per register) are loaded once.
//
// The phase step runs in three passes over the lane's 8 partner pairs: (A) real-FFT split of both bins of every pair - the
// partner's split reuses the products of the first (w[128-k] = -conj(w[k]) to the bit) - which leaves 16 (re, im) values in
// the registers that held the spectrum; (B) exp(angle) of the 16 values, FOUR interleaved per call (exp_angle.cuh);
// (C) Z = S exp(angle) and the inverse split.  Per-block set-up (40 exp() and the 2-tap inverse mel matrix) reads its tables
// from shared memory without branches: it was 17 % of the kernel's stall samples, all of them waiting on dependent loads.
#pragma once

namespace sgs {

constexpr int kG8RowStride = 9;                       // 8 entries + 1 pad: conflict-free row- and column-wise access
constexpr int kG8BufCplx = 16 * kG8RowStride;         // one group's transposition buffer (complex entries)
constexpr int kG8BlockDoubles = 2 * kG8BufCplx * 2;   // x[480] and the two groups' buffers share this space (576 doubles)
constexpr int kG8SLen = 130;
constexpr int kG8MelMax = (kG8BlockDoubles - 480) / 2;   // mel coefficients whose exp() fits behind the waveform (48)

struct G8WarpSmem {
    double xb[2][kG8BlockDoubles];                    // per block: waveform (480 doubles) / transposition buffers
    double S[2][2][kG8SLen];                          // per block, per frame: magnitudes of bins 0..128
};

__device__ __forceinline__ cplx cmulc(cplx a, double wr, double wi) { return {fma(a.x, wr, -a.y * wi), fma(a.x, wi, a.y * wr)}; }

// forward 4-point DFT in place: (a, b, c, d) -> (X0, X1, X2, X3)
__device__ __forceinline__ void dft4(cplx& a, cplx& b, cplx& c, cplx& d) {
    const cplx t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = crot<-1>(csub(b, d));
    a = cadd(t0, t2); b = cadd(t1, t3); c = csub(t0, t2); d = csub(t1, t3);
}

// forward 16-point DFT of v[0..15] (natural order in); X[k] ends at v[pos16(k)]
__device__ __forceinline__ constexpr int pos16(int k) { return 4 * (k & 3) + (k >> 2); }
__device__ __forceinline__ void dft16(cplx (&v)[16]) {
    constexpr double C = 0.92387953251128675613, S = 0.38268343236508977173, H = 0.70710678118654752440;
#pragma unroll
    for (int r = 0; r < 4; ++r) dft4(v[r], v[r + 4], v[r + 8], v[r + 12]);        // v[r + 4s] = u[r][s]
    // u[r][s] *= W16^(r s)
    v[1 + 4] = cmulc(v[1 + 4], C, -S);  v[1 + 8] = cmulc(v[1 + 8], H, -H);  v[1 + 12] = cmulc(v[1 + 12], S, -C);
    v[2 + 4] = cmulc(v[2 + 4], H, -H);  v[2 + 8] = crot<-1>(v[2 + 8]);      v[2 + 12] = cmulc(v[2 + 12], -H, -H);
    v[3 + 4] = cmulc(v[3 + 4], S, -C);  v[3 + 8] = cmulc(v[3 + 8], -H, -H); v[3 + 12] = cmulc(v[3 + 12], -C, S);
#pragma unroll
    for (int s = 0; s < 4; ++s) dft4(v[4 * s], v[4 * s + 1], v[4 * s + 2], v[4 * s + 3]);   // v[4s + t] = X[s + 4t]
}

// forward 8-point DFT of v[0..7] (natural order in); X[k] ends at v[pos8(k)]
__device__ __forceinline__ constexpr int pos8(int k) { return 2 * (k & 3) + (k >> 2); }
__device__ __forceinline__ void dft8(cplx (&v)[8]) {
    constexpr double H = 0.70710678118654752440;
    dft4(v[0], v[2], v[4], v[6]);                                                 // v[2s]     = u0[s]
    dft4(v[1], v[3], v[5], v[7]);                                                 // v[2s + 1] = u1[s]
    v[3] = cmulc(v[3], H, -H);  v[5] = crot<-1>(v[5]);  v[7] = cmulc(v[7], -H, -H);   // u1[s] *= W8^s
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const cplx a = v[2 * s], b = v[2 * s + 1];
        v[2 * s] = cadd(a, b);                                                    // X[s]
        v[2 * s + 1] = csub(a, b);                                                // X[s + 4]
    }
}

struct G8Tables {
    const double* window;      // blackman(256)
    const cplx* tw_t;          // [2][16][9]: W128^((l + 48 f) k1) at [(16 f + k1) * 9 + l]
    const cplx* tw_full;       // exp(-2 pi i k / 256), k <= 128
    const int* inv_idx;
    const double* inv_w;
    int log_mels;              // fromLogMels (exp, non-finite -> 0) or fromMels (as given)
};

// magnitudes are stored as partner pairs (S[p], S[128 - p]) at index p <= 64, so the phase step fetches both with one load
// and carry the inverse transform's 1/256 (a power of two: scaling the magnitudes instead of the 256 output samples changes no bit)
__device__ __forceinline__ void g8_store_mag(double* S, int bin, double v) {
    v *= 1.0 / kFft;
    if (bin <= kHalf / 2) S[2 * bin] = v;
    if (bin >= kHalf / 2) S[2 * (kHalf - bin) + 1] = v;
}

__device__ __forceinline__ cplx csel(bool c, cplx a, cplx b) { return {c ? a.x : b.x, c ? a.y : b.y}; }
__device__ __forceinline__ double shfl_xor_f64(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// synthesis x = irfft(Z0) w at 0 (+) irfft(Z1) w at 160 from v = conj(x~) of slot s at v[pos16(s)]; ANALYSE: multiplied by the
// window again for the next iteration's transform (the window pair of a slot is loaded once for both)
template <bool ANALYSE>
__device__ __forceinline__ void g8_synthesis(cplx (&v)[16], const double2* win_lo, const double2* win_hi) {
    cplx x[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const double2 wv = (s < 10 ? win_lo : win_hi)[8 * s];
        cplx r = cplx{v[pos16(s)].x * wv.x, -v[pos16(s)].y * wv.y};
        // block samples 160..255 sit in slots 10..15 of both frames' lanes with the same l8: x = r0 + r1
        if (s >= 10) r = cplx{r.x + shfl_xor_f64(r.x, 8), r.y + shfl_xor_f64(r.y, 8)};
        x[s] = ANALYSE ? cplx{r.x * wv.x, r.y * wv.y} : r;
    }
#pragma unroll
    for (int s = 0; s < 16; ++s) v[s] = x[s];
}

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_gl_blocks8(const double* __restrict__ logmel, const double* __restrict__ noise, unsigned long long seed,
             double* __restrict__ blocks, const G8Tables tab, int n_frames, int n_mels, int first_frame, int iters,
             long long n_items, long long ring_base, int ring_len) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_window = reinterpret_cast<double*>(smem_raw);                 // [256]
    cplx* s_tw_full = reinterpret_cast<cplx*>(s_window + kFft);             // [130]
    cplx* s_tw_t = s_tw_full + kG8SLen;                                     // [2][16 * 9]
    double2* s_inv_w = reinterpret_cast<double2*>(s_tw_t + 2 * kG8BufCplx); // [130] the two inverse-mel taps of each bin
    int2* s_inv_i = reinterpret_cast<int2*>(s_inv_w + kG8SLen);             // [130] their mel indices
    double* s_ea = reinterpret_cast<double*>(s_inv_i + kG8SLen);            // [520] constants of exp(angle)
    G8WarpSmem* ws_all = reinterpret_cast<G8WarpSmem*>(s_ea + kEaTabLen);
    exp_angle_load_table(s_ea);
    for (int i = threadIdx.x; i < kFft; i += blockDim.x) s_window[i] = tab.window[i];
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
        s_tw_full[i] = tab.tw_full[i];
        s_inv_w[i] = make_double2(tab.inv_w[2 * i], tab.inv_w[2 * i + 1]);
        s_inv_i[i] = make_int2(tab.inv_idx[2 * i], tab.inv_idx[2 * i + 1]);
    }
    for (int i = threadIdx.x; i < 2 * kG8BufCplx; i += blockDim.x) s_tw_t[i] = tab.tw_t[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l8 = lane & 7, grp = lane >> 3, blk = grp >> 1, frm = grp & 1;
    const bool lane0 = l8 == 0, f1 = frm != 0;
    const int row_a = l8, row_b = lane0 ? 8 : 16 - l8;
    G8WarpSmem& ws = ws_all[warp];
    double* em = ws.xb[blk] + kBlk + frm * kG8MelMax;                       // set-up scratch behind the transposition buffers' first 480 doubles
    cplx* buf = reinterpret_cast<cplx*>(ws.xb[blk]) + frm * kG8BufCplx;     // this group's transposition buffer
    double* Sw = ws.S[blk][frm];
    // register slot s holds point n = l8 + 8 ((s + c) & 15), c = 6 for frame 1: window pair / block sample pair of slot s
    const double2* win_lo = reinterpret_cast<const double2*>(s_window) + l8 + (f1 ? 48 : 0);   // s < 10:  win_lo[8 s]
    const double2* win_hi = reinterpret_cast<const double2*>(s_window) + l8 - (f1 ? 80 : 0);   // s >= 10: win_hi[8 s]
    const int pos_lo = frm * kHop + 2 * l8 + (f1 ? 96 : 0), pos_hi = 2 * l8;                  // block sample = pos + 16 s
    const cplx* twt = s_tw_t + frm * kG8BufCplx;                            // W128^((l + 48 frm) k1) at [k1 * 9 + l]
    const int per_sess = n_frames - first_frame;
    const double exp_pi = kExpPi;                                           // exp(angle(-1 + 0j))
    const long long n_pairs = (n_items + 1) >> 1;

    for (long long pr = (long long)blockIdx.x * WARPS + warp; pr < n_pairs; pr += (long long)gridDim.x * WARPS) {
        long long item = 2 * pr + blk;
        const bool valid = item < n_items;
        if (!valid) item = n_items - 1;                                      // odd tail: recompute the last block, do not store
        const int sess = (int)(item / per_sess);
        const int k = first_frame + (int)(item - (long long)sess * per_sess);
        const long long frame = (long long)sess * n_frames + k;

        // Set-up, written as ROLLED loops: straight-line it was 36 KB of code run once per block, which together with the 27 KB
        // iteration loop overflowed the 32 KB instruction cache level (no_instruction stalls: 7 % of the kernel's samples).
        // magnitudes of spectral frame k-1+frm at bins 0..128: exp() once per mel coefficient, then the 2-tap inverse mel matrix
        {
            const double* lm = logmel + (frame - 1 + frm) * n_mels;
            if (n_mels <= kG8MelMax) {
#pragma unroll 1
                for (int m = l8; m < n_mels; m += 8) em[m] = tab.log_mels ? exp(lm[m]) : lm[m];
                __syncwarp();
#pragma unroll 2
                for (int b = l8; b < kBins; b += 8) {
                    const double2 w = s_inv_w[b];
                    const int2 id = s_inv_i[b];
                    const double e0 = em[id.x], e1 = em[id.y];
                    double v = (w.x != 0.0) ? e0 * w.x : 0.0;
                    v = (w.y != 0.0) ? fma(e1, w.y, v) : v;
                    v = (tab.log_mels && !isfinite(v)) ? 0.0 : v;           // MelFilterBank.makeNormal (fromLogMels only)
                    g8_store_mag(Sw, b, v);
                }
            } else {
                for (int b = l8; b < kBins; b += 8) g8_store_mag(Sw, b, mel_magnitude(lm, tab.inv_idx, tab.inv_w, b, tab.log_mels));
            }
        }
        // the initial waveform: into the (still unused) transposition space, then into the register slots
        {
            double* xs = ws.xb[blk];
            const int l16 = lane & 15;
            if (noise) {
                const double2* src = reinterpret_cast<const double2*>(noise + frame * kBlk);
#pragma unroll 1
                for (int i = l16; i < kBlk / 2; i += 16) reinterpret_cast<double2*>(xs)[i] = src[i];
            } else {
#pragma unroll 1
                for (int i = l16; i < kBlk; i += 16) xs[i] = uniform01(seed, (unsigned long long)(ring_base + frame), (unsigned)i);
            }
        }
        __syncwarp();
        cplx v[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int p = (s < 10 ? pos_lo : pos_hi) + 16 * s;
            const double2 xv = *reinterpret_cast<const double2*>(ws.xb[blk] + p);
            const double2 wv = (s < 10 ? win_lo : win_hi)[8 * s];
            v[s] = iters > 0 ? cplx{xv.x * wv.x, xv.y * wv.y} : cplx{xv.x, xv.y};
        }
        __syncwarp();                                                        // magnitudes visible; the scratch is dead

        if (iters > 0) {
#pragma unroll 1
        for (int it = 0;; ++it) {
            // ---- analysis: v = windowed packed frame ----------------------------------------------------------------
            dft16(v);
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) v[pos16(k1)] = cmul(v[pos16(k1)], twt[k1 * kG8RowStride + l8]);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) buf[k1 * kG8RowStride + l8] = v[pos16(k1)];
            __syncwarp();
            cplx ra[8], rb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { ra[j] = buf[row_a * kG8RowStride + j]; rb[j] = buf[row_b * kG8RowStride + j]; }
            dft8(ra);                                                        // bin row_a + 16 k2 at ra[pos8(k2)]
            dft8(rb);
            // ---- phase step on the 8 partner pairs of this lane -----------------------------------------------------
            // lanes >= 1: (row_a, k2) <-> (row_b, 7 - k2).  lane 0: row 8 pairs within itself (k2 <-> 7 - k2), row 0 pairs
            // (k2 <-> 8 - k2) with the self-pair k2 = 4 and DC / Nyquist at k2 = 0.  U[i] / V[7 - i] are the operands of pair i.
            cplx U[8], V[8];
            const cplx dc = ra[pos8(0)];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                U[i] = csel(lane0, rb[pos8(i)], ra[pos8(i)]);
                U[4 + i] = csel(lane0, ra[pos8(1 + i)], ra[pos8(4 + i)]);
                V[i] = csel(lane0, ra[pos8(4 + i)], rb[pos8(i)]);
                V[4 + i] = rb[pos8(4 + i)];
            }
            // (A) split: X[kk] = E + O, X[128 - kk] = conj(E - O): the partner's twiddle w[128 - kk] = -conj(w[kk]) (the host
            // table is built with that symmetry) makes its products the first one's, negated
            double re[16], im[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = lane0 ? (i < 4 ? 8 + 16 * i : 16 * (i - 3)) : l8 + 16 * i;
                const cplx w = s_tw_full[kk];
                const cplx A = U[i], B = V[7 - i];
                // 2 X instead of X: the angle does not see a positive scale, and a factor of two goes through the reciprocal
                // seeds and every product of exp(angle) exactly, so the halves (2 DMUL per pair, 4 DFMA that become DADD) can
                // go without changing a bit of the result
                const cplx d1 = cplx{A.x - B.x, A.y + B.y}, t1 = cmul(w, d1);
                const double ex = A.x + B.x, ey = A.y - B.y;
                re[2 * i] = ex + t1.y;     im[2 * i] = ey - t1.x;
                re[2 * i + 1] = ex - t1.y; im[2 * i + 1] = -ey - t1.x;
            }
            // (B) exp(angle) of the 16 bins, four at a time
            double ea[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const Ea4 e = exp_angle4_call(s_ea, im[4 * g], re[4 * g], im[4 * g + 1], re[4 * g + 1], im[4 * g + 2], re[4 * g + 2],
                                              im[4 * g + 3], re[4 * g + 3]);
                ea[4 * g] = e.a; ea[4 * g + 1] = e.b; ea[4 * g + 2] = e.c; ea[4 * g + 3] = e.d;
            }
            // (C) Z = S exp(angle X) (real, quirk Q1) and the inverse split, stored conjugated
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = lane0 ? (i < 4 ? 8 + 16 * i : 16 * (i - 3)) : l8 + 16 * i;
                const cplx w = s_tw_full[kk];
                const bool upper = !lane0 && i >= 4;                         // kk > 64: the pair sits at 128 - kk, swapped
                const double2 sp = *reinterpret_cast<const double2*>(Sw + 2 * (upper ? kHalf - kk : kk));
                const double z1 = (upper ? sp.y : sp.x) * ea[2 * i], z2 = (upper ? sp.x : sp.y) * ea[2 * i + 1];
                const double sm = z1 + z2, df = z1 - z2;
                const double yi = -(w.x * df);
                U[i] = cplx{fma(w.y, df, sm), yi};
                V[7 - i] = cplx{fma(w.y, -df, sm), yi};                      // pair 7 of lane 0 is the self-pair (64, 64): its V[0] is unused
            }
            cplx zdcny;
            {
                // DC and Nyquist are real with imag = +0.0 in numpy: angle is 0 or pi
                const double xdc = dc.x + dc.y, xny = dc.x - dc.y;
                const double zdc = Sw[0] * ((xdc < 0.0 || (xdc == 0.0 && signbit(xdc))) ? exp_pi : 1.0);
                const double zny = Sw[1] * ((xny < 0.0 || (xny == 0.0 && signbit(xny))) ? exp_pi : 1.0);
                zdcny = cplx{zdc + zny, -(zdc - zny)};
            }
            // back to rows (natural k2 order for the next transform)
            cplx ia[8], ib[8];
            ia[0] = csel(lane0, zdcny, U[0]);
#pragma unroll
            for (int i = 1; i < 4; ++i) ia[i] = csel(lane0, U[3 + i], U[i]);
            ia[4] = csel(lane0, U[7], U[4]);
#pragma unroll
            for (int i = 5; i < 8; ++i) ia[i] = csel(lane0, V[i - 4], U[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { ib[i] = csel(lane0, U[i], V[i]); ib[4 + i] = V[4 + i]; }
            // ---- inverse (forward transform of the conjugated spectrum): 8-point over k2, twiddle, transpose, 16-point over k1
            dft8(ia);                                                        // G[row_a][l] at ia[pos8(l)]
            dft8(ib);
            __syncwarp();                                                    // all row reads of buf are done
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                // l = 0: 1 for frame 0, W16^(6 k1) for frame 1
                buf[row_a * kG8RowStride + l] = cmul(ia[pos8(l)], twt[row_a * kG8RowStride + l]);
                buf[row_b * kG8RowStride + l] = cmul(ib[pos8(l)], twt[row_b * kG8RowStride + l]);
            }
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) v[k1] = buf[k1 * kG8RowStride + l8];
            dft16(v);                                                        // conj(x~) of slot s at v[pos16(s)]
            __syncwarp();                                                    // column reads done before the next stores
            if (it + 1 >= iters) break;
            g8_synthesis<true>(v, win_lo, win_hi);
        }
        g8_synthesis<false>(v, win_lo, win_hi);
        }
        if (valid) {
            // v = the block's samples: frame 0 owns [0, 256), frame 1 [256, 416); [416, 480) = 0 after any iteration
            const long long row = ring_len ? ((ring_base + k) & (ring_len - 1)) : frame;
#pragma unroll
            for (int s = 0; s < 16; ++s)
                if (!f1 || s < 10) *reinterpret_cast<double2*>(blocks + row * kBlk + (s < 10 ? pos_lo : pos_hi) + 16 * s) = make_double2(v[s].x, v[s].y);
            if (f1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int p = kHop + kFft + 2 * l8 + 16 * i;             // 416 + ...
                    double2 t = make_double2(0.0, 0.0);
                    if (iters == 0) {
                        if (noise) t = *reinterpret_cast<const double2*>(noise + frame * kBlk + p);
                        else t = make_double2(uniform01(seed, (unsigned long long)(ring_base + frame), (unsigned)p),
                                              uniform01(seed, (unsigned long long)(ring_base + frame), (unsigned)p + 1));
                    }
                    *reinterpret_cast<double2*>(blocks + row * kBlk + p) = t;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace sgs
