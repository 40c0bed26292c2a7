"""Modal tail of the feature cascade: what the far past of the input leaves in the filter state.

The balanced-pieces feature scan (csrc/feat.cu:k_iir_pieces) starts a time piece from the state the cascade would have had
there.  Round 1 / 2 got it by running all 24 sections from the zero state over the `horizon` samples before the piece
(36 864 at 2048 Hz: the narrowest notch poles have radius 0.9988) - 14 % of the scan's arithmetic.  But only a few modes
live that long.  With s(t+1) = A s(t) + B x[t] and A^n B = sum_m 2 Re(Gamma_m lambda_m^n) (poles are distinct: every section
of a Butterworth band filter has its own complex pair),

    s(t) = sum_m 2 Re( Gamma_m c_m(t) ),      c_m(t) = sum_{k >= 0} lambda_m^k x[t - 1 - k],

and c_m is ONE complex geometric sum per mode - two fused multiply-adds per sample (as a real second-order recurrence)
instead of the cascade's ~99.  So the
warm-up is split at t_near = t0 - near_len: modes that have not decayed to `tol` after near_len samples are summed directly
over their own horizon (csrc/feat.cu:k_iir_tail), the cascade then runs from sum_m 2 Re(Gamma_m c_m) over the last near_len
samples, during which everything that was left out decays below `tol`.

Gamma_m is closed-form per section (no eigen-solver: the 48 x 48 transition matrix is highly non-normal, cond(V) = 5e7):
  * sections before the mode's own section j do not carry it;
  * section j: z' = A_j z + B_j v with A_j = [[-a1, 1], [-a2, 0]], B_j = (b1 - a1 b0, b2 - a2 b0); right / left eigenvectors
    u = (1, -a2 / lambda), l = (1, 1 / lambda) lambda^2 / (lambda^2 - a2); its input carries the mode with amplitude
    G(lambda) = prod_{i < j} H_i(lambda);
  * sections i > j are driven at "frequency" lambda: U = (lambda I - A_i)^-1 B_i Y_in, Y_out = b0 Y_in + U_0.
Evaluated with mpmath at 40 digits (a few milliseconds), rounded to float64 at the end.
"""
import numpy as np

BLOCK = 32            # samples per block of the sum (csrc/feat.cuh:kTailBlock)
MAX_MODES = 16        # csrc/feat.cuh:kTailMaxModes
MODE_GROUP = 4        # modes join the sum in groups of this many (template instantiations of the block routine)
TAIL_WARPS = 16       # csrc/feat.cuh:kTailWarps
CASCADE_OPS = 99.0    # fp64 instructions per sample of the full cascade (24 sections), for the cost model
MODE_OPS = 2.0


def _mp():
    import mpmath
    return mpmath


def modal_expansion(coef, dps=40):
    """coef: (n_biquads, 5) = b0 b1 b2 a1 a2.  Returns (lam[n_biquads] complex128 with imag > 0,
    gamma[2 n_biquads, n_biquads] complex128) with A^n B = sum_m 2 Re(gamma[:, m] lam[m]^n)."""
    mp = _mp()
    nb = len(coef)
    with mp.workdps(dps):
        c = [[mp.mpf(float(v)) for v in row] for row in coef]
        lam = []
        for b0, b1, b2, a1, a2 in c:
            disc = a1 * a1 - 4 * a2
            if disc >= 0:
                raise ValueError("modal tail needs complex pole pairs")
            lam.append(mp.mpc(-a1 / 2, mp.sqrt(-disc) / 2))
        gamma = np.zeros((2 * nb, nb), dtype=np.complex128)
        for j in range(nb):
            z = lam[j]
            G = mp.mpc(1)
            for i in range(j):
                b0, b1, b2, a1, a2 = c[i]
                G *= (b0 * z * z + b1 * z + b2) / (z * z + a1 * z + a2)
            b0, b1, b2, a1, a2 = c[j]
            l0 = z * z / (z * z - a2)
            lB = l0 * (b1 - a1 * b0) + (l0 / z) * (b2 - a2 * b0)
            Y = lB * G                                  # amplitude of lam^n in section j's first state = its output
            col = [mp.mpc(0)] * (2 * nb)
            col[2 * j] = Y
            col[2 * j + 1] = -a2 / z * Y
            for i in range(j + 1, nb):
                b0, b1, b2, a1, a2 = c[i]
                B0, B1 = (b1 - a1 * b0) * Y, (b2 - a2 * b0) * Y
                # (z I - A_i) U = B:  [[z + a1, -1], [a2, z]] U = (B0, B1)
                det = (z + a1) * z + a2
                U0 = (z * B0 + B1) / det
                U1 = ((z + a1) * B1 - a2 * B0) / det
                col[2 * i], col[2 * i + 1] = U0, U1
                Y = b0 * Y + U0
            for i in range(2 * nb):
                gamma[i, j] = complex(col[i])
        lam = np.array([complex(v) for v in lam])
    return lam, gamma


def modal_functionals(coef, sections, dps=40):
    """kappa[m] (complex128, 2 n_biquads) for the pole of each section in `sections`: a state s contributes
    sum_m 2 Re(gamma[:, m] lam[m]^n kappa[m] . s) to the state n samples later - the left eigenvectors of the transition
    matrix, scaled so that kappa[m] . B = 1.  The matrix is block lower triangular (a section only sees the sections before
    it), so w (A - lam I) = 0 is a back-substitution over 2 x 2 blocks from the mode's own section down to the first."""
    mp = _mp()
    nb = len(coef)
    ns = 2 * nb
    with mp.workdps(dps):
        c = [[mp.mpf(float(v)) for v in row] for row in coef]

        def step(s, x):
            s = list(s)
            v = x
            for i in range(nb):
                b0, b1, b2, a1, a2 = c[i]
                o = b0 * v + s[2 * i]
                s[2 * i] = b1 * v - a1 * o + s[2 * i + 1]
                s[2 * i + 1] = b2 * v - a2 * o
                v = o
            return s
        cols = []
        for j in range(ns):
            e = [mp.mpf(0)] * ns
            e[j] = mp.mpf(1)
            cols.append(step(e, mp.mpf(0)))
        A = lambda r, q: cols[q][r]                                        # A[r][q]
        B = step([mp.mpf(0)] * ns, mp.mpf(1))
        out = np.zeros((len(sections), ns), dtype=np.complex128)
        for row, j in enumerate(sections):
            b0, b1, b2, a1, a2 = c[j]
            z = mp.mpc(-a1 / 2, mp.sqrt(4 * a2 - a1 * a1) / 2)
            w = [mp.mpc(0)] * ns
            w[2 * j], w[2 * j + 1] = mp.mpc(1), 1 / z
            for i in range(j - 1, -1, -1):
                r0 = -sum(w[k] * A(k, 2 * i) for k in range(2 * i + 2, 2 * j + 2))
                r1 = -sum(w[k] * A(k, 2 * i + 1) for k in range(2 * i + 2, 2 * j + 2))
                # (w0, w1) [[m00, m01], [m10, m11]] = (r0, r1),  M = A_ii - z I
                m00, m01, m10, m11 = A(2 * i, 2 * i) - z, A(2 * i, 2 * i + 1), A(2 * i + 1, 2 * i), A(2 * i + 1, 2 * i + 1) - z
                det = m00 * m11 - m01 * m10
                w[2 * i] = (r0 * m11 - r1 * m10) / det
                w[2 * i + 1] = (r1 * m00 - r0 * m01) / det
            wb = sum(w[k] * B[k] for k in range(ns))
            for k in range(ns):
                out[row, k] = complex(w[k] / wb)
    return out


def cascade_step_matrix(coef):
    """(A, B) of s(t+1) = A s(t) + B x[t] for the DF2T cascade (scipy.signal.sosfilt's recurrence)."""
    nb = len(coef)
    ns = 2 * nb

    def step(s, x):
        s = s.copy()
        v = x
        for i in range(nb):
            b0, b1, b2, a1, a2 = coef[i]
            o = b0 * v + s[2 * i]
            s[2 * i] = b1 * v - a1 * o + s[2 * i + 1]
            s[2 * i + 1] = b2 * v - a2 * o
            v = o
        return s
    A = np.zeros((ns, ns))
    for j in range(ns):
        e = np.zeros(ns)
        e[j] = 1.0
        A[:, j] = step(e, 0.0)
    return A, step(np.zeros(ns), 1.0)


class ModalTail:
    """The split the device uses: near_len, the modes summed beyond it (slowest first, padded to groups of MODE_GROUP),
    their lengths in samples (multiples of BLOCK, counted back from t_near), the real (2 nb) x (2 n_modes) matrix taking
    (Re c, Im c) to the start state, and the block ranges the kernel's warps sum."""

    def __init__(self, coef, tol, max_horizon=1 << 22):
        coef = np.asarray(coef, dtype=np.float64)
        nb = len(coef)
        lam, gamma = modal_expansion(coef)
        r = np.abs(lam)
        # scale of every state under unit white input: sqrt(sum_n (A^n B)_i^2), by direct recurrence (the modal Gram sum
        # cancels catastrophically: single amplitudes reach 1e6 times the state)
        A, B = cascade_step_matrix(coef)
        v, acc = B.copy(), np.zeros(2 * nb)
        for _ in range(8192):
            acc += v * v
            v = A @ v
        scale = np.sqrt(acc)
        self.state_scale = scale
        # mode m may be dropped n samples back when what it can still contribute (white input beyond n) is below tol of the state
        amp = np.max(2.0 * np.abs(gamma) / scale[:, None], axis=0) / np.sqrt(1.0 - r * r)
        need = np.ceil(np.log(tol / amp) / np.log(r)).astype(np.int64)
        need = np.maximum(need, 0)
        if need.max() > max_horizon:
            raise ValueError("modal tail: horizon %d too long" % need.max())
        order = np.argsort(-need, kind='stable')
        need_sorted = need[order]
        # near_len: multiples of 512 (a segment starts on a 64-sample batch boundary and the state pass wants whole batches);
        # cost = cascade over near_len + 2 flops per sample and live mode beyond it, at most MAX_MODES live modes
        best = None
        for near in range(512, int(need_sorted[0]) + 512, 512):
            live = int(np.sum(need_sorted > near))
            if live > MAX_MODES:
                continue
            groups = -(-live // MODE_GROUP)
            cost = CASCADE_OPS * near
            for g in range(groups):
                L = need_sorted[g * MODE_GROUP] - near
                cost += MODE_OPS * MODE_GROUP * (-(-L // BLOCK) * BLOCK)
            if best is None or cost < best[0]:
                best = (cost, near, live)
        _, near, live = best
        self.cost = best[0]
        self.near_len = int(near)
        n_modes = -(-live // MODE_GROUP) * MODE_GROUP
        self.n_modes = n_modes
        self.lam = np.zeros(n_modes, dtype=np.complex128)
        self.mode_len = np.zeros(n_modes, dtype=np.int32)
        self.gamma = np.zeros((2 * nb, n_modes), dtype=np.complex128)
        for g in range(n_modes // MODE_GROUP):
            L = int(need_sorted[g * MODE_GROUP] - near)
            L = -(-L // BLOCK) * BLOCK
            for q in range(MODE_GROUP):
                k = g * MODE_GROUP + q
                self.mode_len[k] = L
                if k < live:
                    self.lam[k] = lam[order[k]]
                    self.gamma[:, k] = gamma[:, order[k]]
        # what a state leaves in the kept modes (segments nearer to the start of the recording than the horizon still see
        # the reference's initial state): kappa (n_modes x 2 nb, complex; zero rows for padding modes)
        self.kappa = np.zeros((n_modes, 2 * nb), dtype=np.complex128)
        if live:
            self.kappa[:live] = modal_functionals(coef, [int(order[k]) for k in range(live)])
        self.far_len = int(self.mode_len[0]) if n_modes else 0
        self.horizon = self.near_len + self.far_len
        # start state = M (Re c_0.., Im c_0..): 2 Re(gamma c) = 2 Re(gamma) Re(c) - 2 Im(gamma) Im(c)
        self.state_matrix = np.ascontiguousarray(np.concatenate([2.0 * self.gamma.real, -2.0 * self.gamma.imag], axis=1))
        # the (group, block) pairs of all groups laid end to end, dealt to the warps in equal runs; a warp's share of a group
        # is one contiguous range of blocks [lo, hi) (block 0 ends at t_near), and lam^(BLOCK lo) moves its sum to t_near
        n_groups = MAX_MODES // MODE_GROUP
        blocks = [int(self.mode_len[g * MODE_GROUP]) // BLOCK if g * MODE_GROUP < n_modes else 0 for g in range(n_groups)]
        offs = np.concatenate([[0], np.cumsum(blocks)])
        Q = int(offs[-1])
        self.warp_blocks = np.zeros((TAIL_WARPS, n_groups, 2), dtype=np.int32)
        for w in range(TAIL_WARPS):
            q_lo, q_hi = Q * w // TAIL_WARPS, Q * (w + 1) // TAIL_WARPS
            for g in range(n_groups):
                lo, hi = max(q_lo, int(offs[g])), min(q_hi, int(offs[g + 1]))
                if hi > lo:
                    self.warp_blocks[w, g] = (lo - offs[g], hi - offs[g])
        mp = _mp()
        self.warp_shift = np.zeros((TAIL_WARPS, max(n_modes, 1)), dtype=np.complex128)
        with mp.workdps(40):
            for m in range(n_modes):
                z = mp.mpc(self.lam[m].real, self.lam[m].imag)
                for w in range(TAIL_WARPS):
                    self.warp_shift[w, m] = complex(z ** (BLOCK * int(self.warp_blocks[w, m // MODE_GROUP, 0])))

    def reference_state(self, x_far, init=None):
        """Start state at t_near from the samples x_far before it (numpy; for tests).  init: the state before x_far[0]
        when x_far is the whole past (shorter than the horizon), else the past beyond the horizon is taken as forgotten."""
        x_far = np.asarray(x_far, dtype=np.float64)
        c = np.zeros(self.n_modes, dtype=np.complex128)
        for m in range(self.n_modes):
            L = min(int(self.mode_len[m]), len(x_far))
            if L == 0:
                continue
            k = np.arange(L)
            c[m] = np.sum(self.lam[m] ** k * x_far[::-1][:L])
        if init is not None:
            c += self.lam ** len(x_far) * (self.kappa @ np.asarray(init, dtype=np.float64))
        return self.state_matrix @ np.concatenate([c.real, c.imag])

    def kernel_state(self, x_far, init=None):
        """What csrc/feat.cu:k_iir_tail computes, step by step in numpy (for tests): every warp runs the real second-order
        recurrence over its blocks of every group from zero, turns (y(n), y(n-1)) into the complex sum, moves it to t_near,
        the sums are added in warp order and multiplied by the state matrix.  x_far: samples before t_near (oldest first);
        with init (the state before x_far[0]) x_far is the whole past: blocks before it are skipped, the initial state
        enters through kappa and lambda^t_near by repeated squaring."""
        x_far = np.asarray(x_far, dtype=np.float64)
        assert init is None or len(x_far) % BLOCK == 0
        n_groups = MAX_MODES // MODE_GROUP
        tot = np.zeros(self.n_modes, dtype=np.complex128)
        for w in range(TAIL_WARPS):
            for g in range(n_groups):
                lo, hi = (int(v) for v in self.warp_blocks[w, g])
                if init is not None:
                    hi = min(hi, len(x_far) // BLOCK)
                if hi <= lo:
                    continue
                seg = x_far[len(x_far) - BLOCK * hi:len(x_far) - BLOCK * lo]
                for q in range(MODE_GROUP):
                    m = g * MODE_GROUP + q
                    a, b = 2.0 * self.lam[m].real, -(self.lam[m].real ** 2 + self.lam[m].imag ** 2)
                    y1 = y2 = 0.0
                    for v in seg:
                        y1, y2 = a * y1 + (b * y2 + v), y1
                    c = complex(y1 - self.lam[m].real * y2, self.lam[m].imag * y2)
                    tot[m] += c * self.warp_shift[w, m]
        if init is not None:
            for m in range(self.n_modes):
                z, base, n = 1.0 + 0.0j, complex(self.lam[m]), len(x_far)
                while n:
                    if n & 1:
                        z *= base
                    base *= base
                    n >>= 1
                tot[m] += z * np.dot(self.kappa[m], np.asarray(init, dtype=np.float64))
        return self.state_matrix @ np.concatenate([tot.real, tot.imag])


_tails = {}


def modal_tail(coef, tol):
    """ModalTail of a coefficient table, built once per process and configuration (0.1-0.5 s of mpmath)."""
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    key = (coef.tobytes(), float(tol))
    if key not in _tails:
        _tails[key] = ModalTail(coef, tol)
    return _tails[key]
