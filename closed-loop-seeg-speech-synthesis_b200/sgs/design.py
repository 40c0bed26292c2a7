"""Host-side, once-per-configuration tables for the neural->audio path.

Everything here is O(1) in the recording length: filter coefficients, initial filter states,
window start tables, mel (inverse) matrices, smoothing taps. The tables are handed to the CUDA
library (csrc/, include/sgs.h) as plain arrays; no sample data is touched on the host.

Reference behaviour restated (file:line under the reference tree):
  * filter design: livenodes/ECogFeatCalc.py:42-59,126-135 and local/offline.py:24-41 call
    mne.filter.create_filter(..., method='iir', iir_params={'order': 8, 'ftype': 'butter'})["sos"],
    which for that argument set is scipy.signal.iirfilter(8, [f1,f2]/(sr/2), band, 'butter', 'sos').
  * steady-state initial conditions: scipy.signal.sosfilt_zi (FrameBuffer.py:87, offline.py:39-41).
  * warm start of the last notch over (frame_size - shift) zeros: FrameBuffer.py:95-98, offline.py:47-62.
  * frame end positions (online): FrameBuffer.py:27,35,177 ; window starts (offline): offline.py:100-106.
  * mel filter bank: local/MelFilterBank.py:5-83.
  * output low-pass: livenodes/GriffinLim.py:53-59.
"""
import math

import numpy as np
import scipy.signal

N_SECTIONS = 8          # order-8 Butterworth band filters -> 8 biquads each


def band_sos(sr, l_freq, h_freq, order=8):
    """SOS of the order-`order` Butterworth band-pass (l<h) or band-stop (l>h) the reference designs."""
    nyq = sr / 2.0
    if l_freq < h_freq:
        return scipy.signal.iirfilter(order, [l_freq / nyq, h_freq / nyq], btype='bandpass', ftype='butter', output='sos')
    return scipy.signal.iirfilter(order, [h_freq / nyq, l_freq / nyq], btype='bandstop', ftype='butter', output='sos')


def feature_filters(sr, line_noise=50):
    """List of SOS arrays (each (8,6)): high-gamma band-pass then the power-line harmonic notches."""
    filters = [band_sos(sr, 70, 170)]
    if line_noise == 50:
        filters += [band_sos(sr, 102, 98), band_sos(sr, 152, 148)]
    elif line_noise == 60:
        filters += [band_sos(sr, 122, 118)]
    else:
        # the reference builds no notch at all for other values and then fails when wiring the graph
        raise ValueError("line_noise must be 50 or 60")
    return filters


def sosfilt_host(sos, x, zi):
    """Tiny float64 DF2T cascade for O(100)-sample host-side table work (warm-start states).
    Same recurrence as scipy's `_sosfilt`: y=b0*x+z0; z0=b1*x-a1*y+z1; z1=b2*x-a2*y."""
    zi = np.array(zi, dtype=np.float64, copy=True)
    y = np.empty(len(x), dtype=np.float64)
    for n in range(len(x)):
        v = float(x[n])
        for s in range(sos.shape[0]):
            b0, b1, b2, _, a1, a2 = sos[s]
            o = b0 * v + zi[s, 0]
            zi[s, 0] = b1 * v - a1 * o + zi[s, 1]
            zi[s, 1] = b2 * v - a2 * o
            v = o
        y[n] = v
    return y, zi


def propagate(M, n):
    """M^n by n sequential multiplications.  Repeated squaring (numpy.linalg.matrix_power) is numerically
    useless here: these transition matrices are highly non-normal (transient growth 1e3..1e6 before the decay),
    and squaring amplifies the rounding error by |M^(n/2)|^2."""
    X = np.eye(M.shape[0])
    for _ in range(int(n)):
        X = M @ X
    return X


class FeaturePlan:
    """All constants of one feature-extraction configuration (sample rate, window, shift, line noise)."""

    def __init__(self, sr, window_length=None, window_shift=None, line_noise=50, model_order=4, step_size=5,
                 frame_len_ms=None, frame_shift_ms=None):
        # the offline entry point takes seconds (offline.py:12), the node takes milliseconds
        # (ECogFeatCalc.py:19); keep whichever float the caller supplied and derive the other
        if window_length is None:
            window_length = 0.05 if frame_len_ms is None else float(frame_len_ms) / 1000.0
        if window_shift is None:
            window_shift = 0.01 if frame_shift_ms is None else float(frame_shift_ms) / 1000.0
        self.frame_len_ms = float(window_length * 1000.0 if frame_len_ms is None else frame_len_ms)
        self.frame_shift_ms = float(window_shift * 1000.0 if frame_shift_ms is None else frame_shift_ms)
        self.sr = sr
        self.line_noise = line_noise
        self.model_order = int(model_order)
        self.step_size = int(step_size)
        self.window_length = window_length
        self.window_shift = window_shift
        self.filters = feature_filters(sr, line_noise)
        self.n_filters = len(self.filters)
        self.n_biquads = self.n_filters * N_SECTIONS
        # coefficient table, one row per biquad: b0 b1 b2 a1 a2 (a0 == 1)
        sos = np.vstack(self.filters)
        assert np.all(sos[:, 3] == 1.0)
        self.coef = np.ascontiguousarray(sos[:, [0, 1, 2, 4, 5]], dtype=np.float64)
        # unit steady-state initial conditions per filter, (n_biquads, 2)
        self.zi_unit = np.vstack([scipy.signal.sosfilt_zi(f) for f in self.filters]).astype(np.float64)
        # frame geometry
        self.frame_size = int((self.frame_len_ms / 1000.0) * float(sr))               # 51 @1024, 102 @2048
        self.zero_fill = self.frame_size - int((self.frame_shift_ms / 1000.0) * float(sr))  # 41 / 82
        assert self.frame_size == int(window_length * sr) and self.zero_fill > 0
        # last filter: unit zi advanced over `zero_fill` zeros; the transient it emits is part of
        # the first online frames (FrameBuffer.py:95-98) and is discarded offline (offline.py:62)
        last = self.filters[-1]
        zf_out, zf_state = sosfilt_host(last, np.zeros(self.zero_fill), scipy.signal.sosfilt_zi(last))
        self.zero_fill_response = zf_out
        self.zi_last_warm = zf_state

    # ---- window tables -------------------------------------------------------------------
    def offline_num_windows(self, n_samples):
        """offline.py:100"""
        if n_samples < self.window_length * self.sr:
            return 0
        return int(np.floor((n_samples - self.window_length * self.sr) / (self.window_shift * self.sr))) + 1

    def offline_window_starts(self, n_samples):
        """offline.py:105-106; returns (starts int32[W], length) in real-sample coordinates."""
        nw = self.offline_num_windows(n_samples)
        # int(round(k * shift * sr)) for every k at once: np.rint rounds half to even like Python's round(), and the float64
        # products are the ones the reference's loop forms (an hour of recording has 360 000 windows - as a Python loop this
        # table cost more than the feature kernels)
        k = np.arange(nw, dtype=np.float64)
        starts = np.rint((k * self.window_shift) * self.sr).astype(np.int64)
        ends = np.rint(starts + self.window_length * self.sr).astype(np.int64)
        if nw:
            length = int(ends[0] - starts[0])
            assert int((ends - starts).min()) == length == int((ends - starts).max())
        else:
            length = int(round(self.window_length * self.sr))
        return starts, length

    def online_frame_ends(self, n_samples):
        """FrameBuffer.py:27,35,177: end position (exclusive) of frame k in the zero-fill-prefixed stream.
        Frame k exists once `zero_fill + n_samples >= E_k`."""
        sr = float(self.sr)
        frame_size = self.frame_size
        first_ms = (float(frame_size) / sr) * 1000.0
        shift_ms = self.frame_shift_ms
        total = self.zero_fill + n_samples
        # E_k = round(((first_ms + k * shift_ms) / 1000) * sr), k = 0, 1, ... while E_k <= total - vectorised (np.rint = Python's
        # banker's round; the same float64 expression per k), with the count found from an over-estimate
        n_guess = int(max(0.0, (total - frame_size) / (float(shift_ms) / 1000.0 * sr))) + 3
        k = np.arange(n_guess, dtype=np.float64)
        ends = np.rint(((first_ms + k * shift_ms) / 1000.0) * sr).astype(np.int64)
        ends[0] = frame_size
        stop = int(np.argmax(ends > total)) if (ends > total).any() else len(ends)
        assert stop < n_guess
        return ends[:stop]

    def online_window_starts(self, n_samples):
        """Online frames as (starts, length) in real-sample coordinates (negative = zero-fill region)."""
        ends = self.online_frame_ends(n_samples)
        return ends - self.frame_size - self.zero_fill, self.frame_size

    @property
    def context(self):
        return self.model_order * self.step_size


# ------------------------------------------------------------------------------------------
# mel filter bank (local/MelFilterBank.py:5-83)
# ------------------------------------------------------------------------------------------
class MelTables:
    def __init__(self, spec_size, num_coefficients, sample_rate):
        nb = int(num_coefficients)
        max_mel = 2595.0 * math.log10(1.0 + (sample_rate / 2.0) / 700.0)
        step = (max_mel - 0) / (nb + 1)
        edges = np.arange(0, nb + 2) * step

        def to_bin(m):
            f = math.floor(700.0 * (math.pow(10.0, m / 2595.0) - 1.0))
            return int(math.floor((f / (sample_rate / 2.0)) * spec_size))

        idx = [to_bin(m) for m in edges]
        tri = np.zeros((nb, spec_size))
        for i in range(nb):
            a, c, e = idx[i:i + 3]
            if c > a:
                tri[i, a:c] = (np.arange(a, c) - a) / float(c - a)
            if e > c:
                tri[i, c:e] = (e - np.arange(c, e)) / float(e - c)
        # no copies between the transposes: numpy's pairwise summation order depends on the memory layout, and the
        # reference sums the transposed VIEWS (MelFilterBank.py:35-39); a copy changes the column sums by an ulp
        fwd = tri.transpose()                                # (spec_size, nb)
        fwd = _finite(fwd / _colsum(fwd))
        inv = fwd.transpose()                                # (nb, spec_size)
        inv = _finite(inv / _colsum(inv))
        self.spec_size = spec_size
        self.n_mels = nb
        self.mel = fwd
        self.inv = inv
        # the inverse has at most two non-zeros per spectral bin: store it as a 2-tap table
        tap_idx = np.zeros((spec_size, 2), dtype=np.int32)
        tap_w = np.zeros((spec_size, 2), dtype=np.float64)
        for f in range(spec_size):
            nz = np.nonzero(inv[:, f])[0]
            assert len(nz) <= 2, "inverse mel matrix is expected to be a 2-tap interpolation"
            for j, m in enumerate(nz):
                tap_idx[f, j] = m
                tap_w[f, j] = inv[m, f]
        self.inv_idx = tap_idx
        self.inv_w = tap_w


def _colsum(x):
    s = np.sum(x, axis=0)
    s[s == 0] = 1.0
    return s


def _finite(x):
    x[~np.isfinite(x)] = 0
    return x


# ------------------------------------------------------------------------------------------
# dequantisation smoothing (scipy.ndimage.gaussian_filter(sigma=0.5), Dequantization.py:17)
# ------------------------------------------------------------------------------------------
def gaussian_taps(sigma=0.5, truncate=4.0):
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return w / w.sum()


# ------------------------------------------------------------------------------------------
# Griffin-Lim node constants (livenodes/GriffinLim.py:13-62)
# ------------------------------------------------------------------------------------------
class GriffinLimNodePlan:
    def __init__(self, frame_size_ms=16, frame_shift_ms=10, sample_rate=16000, n_mels=40, iterations=8,
                 extra_context=0, cutoff=7900, norm_factor=1.0):
        fs_ms, sh_ms, sr = float(frame_size_ms), float(frame_shift_ms), float(sample_rate)
        self.sample_rate = sr
        self.frame_shift_ms = sh_ms
        self.fft_size = int((fs_ms / 1000.0) * sr)
        self.hop = int((sh_ms / 1000.0) * sr)
        self.context_width = int(fs_ms / sh_ms)
        self.block_len = self.context_width * 2 + 1 + extra_context
        self.block_samples = self.block_len * self.hop
        self.spec_frames = self.block_len - self.context_width       # frames handed to the block synthesis
        self.iterations = int(iterations)
        self.norm_factor = norm_factor
        self.window = np.blackman(self.fft_size)
        self.ola_window = np.blackman(self.block_samples)
        order = int((sr / 1000.0) * sh_ms / 32.0)
        self.lp_b, self.lp_a = scipy.signal.iirfilter(order, float(cutoff) / float(sr / 2), btype="lowpass")
        self.spec_size = int(self.fft_size / 2 + 1)
        self.mel = MelTables(self.spec_size, n_mels, sr)
        # STFT/ISTFT frame offsets inside a block: range(0, block_samples - fft_size, hop) (quirk Q3)
        self.offsets = list(range(0, self.block_samples - self.fft_size, self.hop))
