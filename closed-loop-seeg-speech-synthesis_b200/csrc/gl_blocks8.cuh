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
// The transposition buffer aliases the block's waveform buffer, which is dead between analysis and synthesis.
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
    const cplx* tw_t;          // [16][9]: W128^(l k1) at [k1 * 9 + l]
    const cplx* tw_full;       // exp(-2 pi i k / 256), k <= 128
    const int* inv_idx;
    const double* inv_w;
};

// one real-FFT bin pair (kk, 128 - kk): split, Z = S exp(angle X), inverse split (stored conjugated); identical
// arithmetic to the shared-memory kernel
__device__ __forceinline__ void g8_pair(const double* s_ea, cplx& A, cplx& B, cplx w, cplx w2, double s1, double s2) {
    const cplx d1 = cplx{A.x - B.x, A.y + B.y}, t1 = cmul(w, d1);
    const double re1 = 0.5 * (A.x + B.x) + 0.5 * t1.y, im1 = 0.5 * (A.y - B.y) - 0.5 * t1.x;
    const cplx d2 = cplx{B.x - A.x, B.y + A.y}, t2 = cmul(w2, d2);
    const double re2 = 0.5 * (B.x + A.x) + 0.5 * t2.y, im2 = 0.5 * (B.y - A.y) - 0.5 * t2.x;
    const double2 ea = exp_angle_pair_call(s_ea, im1, re1, im2, re2);
    const double z1 = s1 * ea.x, z2 = s2 * ea.y;
    const double sm = z1 + z2, df = z1 - z2;
    A = cplx{fma(w.y, df, sm), -(w.x * df)};
    B = cplx{fma(w2.y, -df, sm), w2.x * df};
}

// magnitudes are stored as partner pairs (S[p], S[128 - p]) at index p <= 64, so the phase step fetches both with one load
// and carry the inverse transform's 1/256 (a power of two: scaling the magnitudes instead of the 256 output samples changes no bit)
__device__ __forceinline__ void g8_store_mag(double* S, int bin, double v) {
    v *= 1.0 / kFft;
    if (bin <= kHalf / 2) S[2 * bin] = v;
    if (bin >= kHalf / 2) S[2 * (kHalf - bin) + 1] = v;
}

__device__ __forceinline__ cplx csel(bool c, cplx a, cplx b) { return {c ? a.x : b.x, c ? a.y : b.y}; }

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_gl_blocks8(const double* __restrict__ logmel, const double* __restrict__ noise, unsigned long long seed,
             double* __restrict__ blocks, const G8Tables tab, int n_frames, int n_mels, int first_frame, int iters,
             long long n_items, long long ring_base, int ring_len) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_window = reinterpret_cast<double*>(smem_raw);                 // [256]
    cplx* s_tw_full = reinterpret_cast<cplx*>(s_window + kFft);             // [130]
    cplx* s_tw_t = s_tw_full + kG8SLen;                                     // [16 * 9]
    double* s_ea = reinterpret_cast<double*>(s_tw_t + kG8BufCplx);          // [72] constants of exp(angle)
    G8WarpSmem* ws_all = reinterpret_cast<G8WarpSmem*>(s_ea + kEaTabLen);
    exp_angle_load_table(s_ea);
    for (int i = threadIdx.x; i < kFft; i += blockDim.x) s_window[i] = tab.window[i];
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) s_tw_full[i] = tab.tw_full[i];
    for (int i = threadIdx.x; i < kG8BufCplx; i += blockDim.x) s_tw_t[i] = tab.tw_t[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l8 = lane & 7, grp = lane >> 3, blk = grp >> 1, frm = grp & 1;
    const bool lane0 = l8 == 0;
    const int row_a = l8, row_b = lane0 ? 8 : 16 - l8;
    G8WarpSmem& ws = ws_all[warp];
    double* x = ws.xb[blk];                                                 // this block's waveform
    cplx* buf = reinterpret_cast<cplx*>(ws.xb[blk]) + frm * kG8BufCplx;     // this group's transposition buffer (aliases x)
    const double* Sg = ws.S[blk][frm];
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

        // magnitudes of spectral frame k-1+frm at bins 0..128, and the initial waveform
        {
            const double* lm = logmel + (frame - 1 + frm) * n_mels;
            if (n_mels <= kG8MelMax) {
                // exp() once per mel coefficient (40 per frame) instead of once per inverse-mel tap (258 per frame): the values
                // sit in the tail of the block's buffer, which the 480-sample waveform does not use
                double* em = x + kBlk + frm * kG8MelMax;
                for (int m = l8; m < n_mels; m += 8) em[m] = exp(lm[m]);
                __syncwarp();
                for (int b = l8; b < kBins; b += 8) {
                    const double w0 = tab.inv_w[b * 2], w1 = tab.inv_w[b * 2 + 1];
                    double v = 0.0;
                    if (w0 != 0.0) v = em[tab.inv_idx[b * 2]] * w0;
                    if (w1 != 0.0) v = fma(em[tab.inv_idx[b * 2 + 1]], w1, v);
                    v = isfinite(v) ? v : 0.0;                              // MelFilterBank.makeNormal
                    g8_store_mag(ws.S[blk][frm], b, v);
                }
            } else {
                for (int b = l8; b < kBins; b += 8) g8_store_mag(ws.S[blk][frm], b, mel_magnitude(lm, tab.inv_idx, tab.inv_w, b));
            }
            const int l16 = lane & 15;
            for (int i = l16; i < kBlk; i += 16)
                x[i] = noise ? noise[frame * kBlk + i] : uniform01(seed, (unsigned long long)(ring_base + frame), (unsigned)i);
        }
        __syncwarp();

#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            cplx v[16];
            // ---- analysis: window + pack (z[n] = x[o+2n] w[2n] + i x[o+2n+1] w[2n+1]), n = l8 + 8m ---------------
            {
                const double* xo = x + frm * kHop;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int n = l8 + 8 * m;
                    const double2 xv = *reinterpret_cast<const double2*>(xo + 2 * n);
                    const double2 wv = *reinterpret_cast<const double2*>(s_window + 2 * n);
                    v[m] = cplx{xv.x * wv.x, xv.y * wv.y};
                }
            }
            dft16(v);
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) v[pos16(k1)] = cmul(v[pos16(k1)], s_tw_t[k1 * kG8RowStride + l8]);
            __syncwarp();                                                    // every lane has read x: its space becomes buf
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
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = lane0 ? (i < 4 ? 8 + 16 * i : 16 * (i - 3)) : l8 + 16 * i;
                cplx B = V[7 - i];
                // partner twiddle w[128 - kk] = -conj(w[kk]) (the host table is built with that symmetry)
                const cplx w = s_tw_full[kk];
                const bool upper = !lane0 && i >= 4;                         // kk > 64: the pair sits at 128 - kk, swapped
                const double2 sp = *reinterpret_cast<const double2*>(Sg + 2 * (upper ? kHalf - kk : kk));
                g8_pair(s_ea, U[i], B, w, cplx{0.0 - w.x, w.y}, upper ? sp.y : sp.x, upper ? sp.x : sp.y);      // 0.0 - x: no -0.0 at the quadrant point
                if (i != 7) V[7 - i] = B;                                    // pair 7 of lane 0 is the self-pair (64, 64): V[0] unused
                else V[0] = lane0 ? V[0] : B;
            }
            cplx zdcny;
            {
                // DC and Nyquist are real with imag = +0.0 in numpy: angle is 0 or pi
                const double xdc = dc.x + dc.y, xny = dc.x - dc.y;
                const double zdc = Sg[0] * ((xdc < 0.0 || (xdc == 0.0 && signbit(xdc))) ? exp_pi : 1.0);
                const double zny = Sg[1] * ((xny < 0.0 || (xny == 0.0 && signbit(xny))) ? exp_pi : 1.0);
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
                const cplx ga = l == 0 ? ia[pos8(0)] : cmul(ia[pos8(l)], s_tw_t[row_a * kG8RowStride + l]);
                const cplx gb = l == 0 ? ib[pos8(0)] : cmul(ib[pos8(l)], s_tw_t[row_b * kG8RowStride + l]);
                buf[row_a * kG8RowStride + l] = ga;
                buf[row_b * kG8RowStride + l] = gb;
            }
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) v[k1] = buf[k1 * kG8RowStride + l8];
            dft16(v);                                                        // conj(x~[l8 + 8m]) at v[pos16(m)]
            __syncwarp();                                                    // buf is dead: the space becomes x again
            // ---- synthesis: x = irfft(Z0) w at 0  (+)  irfft(Z1) w at 160; nothing reaches [416, 480) ------------------
            if (frm == 0) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int n = l8 + 8 * m;
                    const double2 wv = *reinterpret_cast<const double2*>(s_window + 2 * n);
                    *reinterpret_cast<double2*>(x + 2 * n) = make_double2(v[pos16(m)].x * wv.x, -v[pos16(m)].y * wv.y);
                }
            }
            __syncwarp();
            if (frm == 1) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {                               // frame 0 first: the overlap is (0 + r0) + r1
                    const int n = l8 + 8 * m, p = 2 * n;
                    const double2 wv = *reinterpret_cast<const double2*>(s_window + p);
                    const double r0 = v[pos16(m)].x * wv.x, r1 = -v[pos16(m)].y * wv.y;
                    double2 cur = *reinterpret_cast<const double2*>(x + kHop + p);
                    cur.x = (p < kFft - kHop) ? cur.x + r0 : r0;
                    cur.y = (p + 1 < kFft - kHop) ? cur.y + r1 : r1;
                    *reinterpret_cast<double2*>(x + kHop + p) = cur;
                }
#pragma unroll
                for (int i = 0; i < (kBlk - kHop - kFft) / 8; ++i) x[kHop + kFft + l8 + 8 * i] = 0.0;
            }
            __syncwarp();
        }
        if (valid) {
            const long long row = ring_len ? ((ring_base + k) & (ring_len - 1)) : frame;
            const int l16 = lane & 15;
            for (int i = l16; i < kBlk; i += 16) blocks[row * kBlk + i] = x[i];
        }
        __syncwarp();
    }
}

}  // namespace sgs
