// Warp-level FP64 FFT building blocks (Stockham autosort, mixed radix 2/4/5) on shared memory.
// One warp transforms one sequence; buffers are warp-private, so stages are separated by __syncwarp only.
// Replaces numpy.fft.rfft / irfft (pocketfft) at livenodes/GriffinLim.py:66,73 and local/offline.py:149,159,236.
#pragma once
#include <cuda_runtime.h>

namespace sgs {

struct cplx { double x, y; };

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.x, -a.y}; }
// multiply by -i (SIGN < 0, forward) or +i (SIGN > 0, inverse)
template <int SIGN> __device__ __forceinline__ cplx crot(cplx a) { return SIGN < 0 ? cplx{a.y, -a.x} : cplx{-a.y, a.x}; }

template <int R, int SIGN> struct Butterfly;

template <int SIGN> struct Butterfly<2, SIGN> {
    __device__ static __forceinline__ void run(cplx (&v)[2]) {
        const cplx a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int SIGN> struct Butterfly<4, SIGN> {
    __device__ static __forceinline__ void run(cplx (&v)[4]) {
        const cplx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        const cplx t2 = cadd(v[1], v[3]), t3 = crot<SIGN>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

template <int SIGN> struct Butterfly<5, SIGN> {
    __device__ static __forceinline__ void run(cplx (&v)[5]) {
        constexpr double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos(2pi/5), cos(4pi/5)
        constexpr double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin(2pi/5), sin(4pi/5)
        const cplx a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
        const cplx b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
        const cplx m1 = {fma(c2, a2.x, fma(c1, a1.x, v[0].x)), fma(c2, a2.y, fma(c1, a1.y, v[0].y))};
        const cplx m2 = {fma(c1, a2.x, fma(c2, a1.x, v[0].x)), fma(c1, a2.y, fma(c2, a1.y, v[0].y))};
        const cplx n1 = crot<SIGN>(cplx{fma(s2, b2.x, s1 * b1.x), fma(s2, b2.y, s1 * b1.y)});
        const cplx n2 = crot<SIGN>(cplx{fma(-s1, b2.x, s2 * b1.x), fma(-s1, b2.y, s2 * b1.y)});
        v[0] = cadd(v[0], cadd(a1, a2));
        v[1] = cadd(m1, n1);
        v[4] = csub(m1, n1);
        v[2] = cadd(m2, n2);
        v[3] = csub(m2, n2);
    }
};

// One Stockham stage of an N-point transform: radix R, NS = product of the radices already applied.
// tw[t] = exp(-2*pi*i*t/N), t < N.  Reads a, writes b.
// STAGED: tw is this stage's own table, tw[(r - 1) * NS + k] = exp(-2*pi*i*r*k/(NS*R)), so that the lanes of a warp
// (consecutive k) read consecutive entries; the shared N-entry table is read with strides of N/(NS*R) entries, which
// for NS = 25, N = 400 puts up to 25 lanes on one bank group.
template <int N, int R, int NS, int SIGN, bool STAGED = false>
__device__ __forceinline__ void stockham_stage(const cplx* __restrict__ a, cplx* __restrict__ b,
                                               const cplx* __restrict__ tw, int lane) {
    constexpr int M = N / R;
    for (int j = lane; j < M; j += 32) {
        const int k = j % NS;
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = a[j + r * M];
        if (NS > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                cplx w = STAGED ? tw[(r - 1) * NS + k] : tw[r * k * (N / (NS * R))];
                if (SIGN > 0) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        Butterfly<R, SIGN>::run(v);
        const int j0 = (j / NS) * NS * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) b[j0 + q * NS] = v[q];
    }
    __syncwarp();
}

// 128-point complex FFT (radices 4,4,4,2): result ends in `a`.
template <int SIGN>
__device__ __forceinline__ void fft128(cplx* a, cplx* b, const cplx* tw, int lane) {
    stockham_stage<128, 4, 1, SIGN>(a, b, tw, lane);
    stockham_stage<128, 4, 4, SIGN>(b, a, tw, lane);
    stockham_stage<128, 4, 16, SIGN>(a, b, tw, lane);
    stockham_stage<128, 2, 64, SIGN>(b, a, tw, lane);
}

// 400-point complex FFT (radices 5,5,4,4): result ends in `a`.  The odd radices go first: a stage writes with a lane
// stride of its radix (NS = 1) or in runs of NS entries, and strides of 5 and 25 complex values (80 / 400 B) spread over
// all banks where 4 and 16 (64 / 256 B) pile 16 lanes onto the same four.
constexpr int kFft400TwLen = 4 * 5 + 3 * 25 + 3 * 100;                     // per-stage twiddle tables of fft400 (395 entries)

// builds the per-stage tables from tw400[t] = exp(-2*pi*i*t/400) (all threads of the CTA; sync afterwards)
__device__ __forceinline__ void fft400_stage_tables(cplx* s_tw, const cplx* __restrict__ tw400) {
    for (int i = threadIdx.x; i < kFft400TwLen; i += blockDim.x) {
        int r, k, step;
        if (i < 20) { r = i / 5 + 1; k = i % 5; step = 16; }                   // NS = 5, R = 5: 400 / 25
        else if (i < 95) { r = (i - 20) / 25 + 1; k = (i - 20) % 25; step = 4; }   // NS = 25, R = 4: 400 / 100
        else { r = (i - 95) / 100 + 1; k = (i - 95) % 100; step = 1; }         // NS = 100, R = 4
        s_tw[i] = tw400[r * k * step];
    }
}

template <int SIGN>
__device__ __forceinline__ void fft400(cplx* a, cplx* b, const cplx* tw_stage, int lane) {
    stockham_stage<400, 5, 1, SIGN>(a, b, tw_stage, lane);
    stockham_stage<400, 5, 5, SIGN, true>(b, a, tw_stage, lane);
    stockham_stage<400, 4, 25, SIGN, true>(a, b, tw_stage + 20, lane);
    stockham_stage<400, 4, 100, SIGN, true>(b, a, tw_stage + 95, lane);
}

template <int M, int SIGN> struct HalfFFT;
template <int SIGN> struct HalfFFT<128, SIGN> { __device__ static __forceinline__ void run(cplx* a, cplx* b, const cplx* tw, int lane) { fft128<SIGN>(a, b, tw, lane); } };
template <int SIGN> struct HalfFFT<400, SIGN> { __device__ static __forceinline__ void run(cplx* a, cplx* b, const cplx* tw, int lane) { fft400<SIGN>(a, b, tw, lane); } };

}  // namespace sgs
