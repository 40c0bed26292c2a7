"""CPU oracle for the neural->audio hot path: a numpy/scipy restatement of the reference algorithms.

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module, and only as the checker or the timed CPU baseline.
The product path (closed-loop-seeg-speech-synthesis_b200/) never imports it and has no CPU fallback.

Pinning: the reference repository ships no tests, fixtures or golden vectors, so nothing upstream
pins these functions.  They are pinned instead against the reference itself executed in the build
container (oracle/gen_golden.py imports /root/reference through oracle/ref_shim.py and writes
tests/golden/*.npz; tests/test_oracle_golden.py checks every function below against those files).
One boundary stays PARITY UNPINNED: filter design goes through `mne.filter.create_filter`
(mne 0.19.0, absent here); it is restated as scipy.signal.iirfilter per mne's documented
behaviour for iir_params={'order': 8, 'ftype': 'butter'}.  Third-party arithmetic the reference
calls (scipy.signal.sosfilt/sosfilt_zi/lfilter/iirfilter, numpy.fft, scipy.ndimage.gaussian_filter,
scipy.stats.spearmanr, sklearn LinearDiscriminantAnalysis) is called here the same way.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
import math

import numpy as np
import scipy.signal
from scipy.ndimage import gaussian_filter
from scipy.signal.windows import hann


# --------------------------------------------------------------------------------------------
# a1: filter design  (livenodes/ECogFeatCalc.py:42-59,126-135 ; local/offline.py:24-37,72-74)
# --------------------------------------------------------------------------------------------
def create_filter_sos(sr, l_freq, h_freq):
    nyq = sr / 2.0
    if l_freq < h_freq:
        return scipy.signal.iirfilter(8, [l_freq / nyq, h_freq / nyq], btype='bandpass', ftype='butter', output='sos')
    return scipy.signal.iirfilter(8, [h_freq / nyq, l_freq / nyq], btype='bandstop', ftype='butter', output='sos')


def _tile_zi(sos, n_ch):
    zi = scipy.signal.sosfilt_zi(sos)                                   # (8, 2)
    return np.repeat(zi, n_ch, axis=-1).reshape(zi.shape[0], zi.shape[1], n_ch)


# --------------------------------------------------------------------------------------------
# a7: filtered stream shared by both feature paths (R1 in SURVEY.md 8a')
#     offline.py:31-97 ; FrameBuffer.py:86-98,139-143 chained as in ECogFeatCalc.py:67-85
# --------------------------------------------------------------------------------------------
def high_gamma_stream(eeg, sr, line_noise=50, window_length=0.05, window_shift=0.01, keep_zero_fill=False, warm_start=True):
    """Returns the notch-filtered high-gamma signal.  With keep_zero_fill the response of the last
    filter to its warm-start zeros is kept in front (the online node frames over it).  warm_start=False
    (ECogFeatCalc(warm_start=False), FrameBuffer.py:91-93): the last filter starts cold like the others, no zero fill."""
    data = np.asarray(eeg, dtype=np.float64)
    n_ch = data.shape[1]
    hg = create_filter_sos(sr, 70, 170)
    if line_noise == 50:
        notches = [create_filter_sos(sr, 102, 98), create_filter_sos(sr, 152, 148)]
    else:
        notches = [create_filter_sos(sr, 122, 118)]
    zero_fill = np.zeros([int(window_length * sr) - int(window_shift * sr), n_ch])

    hg_state = _tile_zi(hg, n_ch) * data[0, :]                           # cold start (offline.py:51-52)
    data, _ = scipy.signal.sosfilt(hg, data, axis=0, zi=hg_state)
    for f in notches[:-1]:                                               # cold start on filtered x[0] (58-59)
        st = _tile_zi(f, n_ch) * data[0, :]
        data, _ = scipy.signal.sosfilt(f, data, axis=0, zi=st)
    last = notches[-1]                                                   # warm start (62 / 93)
    if not warm_start:
        st = _tile_zi(last, n_ch) * data[0, :]
        data, _ = scipy.signal.sosfilt(last, data, axis=0, zi=st)
        return (data, 0) if keep_zero_fill else data
    st = _tile_zi(last, n_ch)
    head, st = scipy.signal.sosfilt(last, zero_fill, axis=0, zi=st)
    data, _ = scipy.signal.sosfilt(last, data, axis=0, zi=st)
    if keep_zero_fill:
        return np.vstack([head, data]), zero_fill.shape[0]
    return data


# --------------------------------------------------------------------------------------------
# a5/a6/a7: offline features  (local/offline.py:12-128)
# --------------------------------------------------------------------------------------------
def herff2016_b(eeg, sr, window_length=0.05, window_shift=0.01, line_noise=50, skip_stacking=False,
                model_order=4, step_size=5):
    data = high_gamma_stream(eeg, sr, line_noise, window_length, window_shift)
    num_windows = int(np.floor((data.shape[0] - window_length * sr) / (window_shift * sr))) + 1
    feat = np.zeros((num_windows, data.shape[1]))
    for win in range(num_windows):                                       # offline.py:104-108
        start = int(round((win * window_shift) * sr))
        stop = int(round(start + window_length * sr))
        for c in range(data.shape[1]):
            feat[win, c] = np.log(np.sum(data[start:stop, c] ** 2) + 0.01)
    if skip_stacking:
        return feat
    return stack_offline(feat, model_order, step_size)


def stack_offline(features, model_order=4, step_size=5):
    """offline.py:111-116"""
    ctx = model_order * step_size
    out = np.zeros([features.shape[0] - ctx, (model_order + 1) * features.shape[1]])
    for f_num, i in enumerate(range(ctx, features.shape[0])):
        out[f_num, :] = features[i - ctx:i + 1:step_size, :].T.flatten()
    return out


# --------------------------------------------------------------------------------------------
# a2-a6: streaming features, closed form of the ECogFeatCalc node chain (R1 "ONLINE")
#     FrameBuffer.py:60-177, ECogFeatCalc.py:67-104,118-144
# --------------------------------------------------------------------------------------------
def ecog_feat_calc(eeg, sr, frame_len_ms=50, frame_shift_ms=10, model_order=4, step_size=5, line_noise=50,
                   chunk_size=32, stacked=True, warm_start=True):
    """All frames the node chain emits after the whole array has been pushed through it (any
    chunking; the chain is chunk-size invariant for chunks < 2048 samples).
    The first two FrameBuffers forward only complete `chunk_size` blocks, so a trailing partial
    block never reaches the last filter."""
    eeg = np.asarray(eeg, dtype=np.float64)
    sr_f = float(sr)
    usable = (eeg.shape[0] // chunk_size) * chunk_size
    n_ch = eeg.shape[1]
    if usable == 0:
        return np.zeros((0, (model_order + 1) * n_ch if stacked else n_ch))
    y2, _ = high_gamma_stream(eeg[:usable], sr, line_noise, frame_len_ms / 1000.0, frame_shift_ms / 1000.0,
                              keep_zero_fill=True, warm_start=warm_start)
    frame_size = int((float(frame_len_ms) / 1000.0) * sr_f)              # FrameBuffer.py:27
    first_ms = (float(frame_size) / sr_f) * 1000.0                       # FrameBuffer.py:35
    feats = []
    k, end = 0, frame_size
    while end <= y2.shape[0]:
        w = y2[end - frame_size:end]
        feats.append(np.log(np.sum(w ** 2, axis=0) + 0.01))              # ECogFeatCalc.py:118-124
        k += 1
        end = round(((first_ms + k * float(frame_shift_ms)) / 1000.0) * sr_f)   # FrameBuffer.py:177
    f = np.array(feats).reshape(-1, n_ch)
    if not stacked:
        return f
    if not warm_start:
        # the stack FrameBuffer (21 rows / 1 row, FrameBuffer.py:99-100) starts empty: first output with the 21st row
        return stack_offline(f, model_order, step_size) if len(f) > model_order * step_size else np.zeros((0, (model_order + 1) * n_ch))
    return stack_online(f, model_order, step_size)


def stack_online(f, model_order=4, step_size=5):
    """Stack FrameBuffer (21 ms / 1 ms @ "1000 Hz", warm start => 20 zero rows) + stack_features
    (ECogFeatCalc.py:99-100,137-144): row k = [f[k-20], f[k-15], ..., f[k]] with zeros before the start."""
    ctx = model_order * step_size
    padded = np.vstack([np.zeros((ctx, f.shape[1])), f])
    out = np.zeros((f.shape[0], (model_order + 1) * f.shape[1]))
    for k in range(f.shape[0]):
        out[k] = padded[k:k + ctx + 1:step_size].T.flatten()
    return out


# --------------------------------------------------------------------------------------------
# a10: mel filter bank  (local/MelFilterBank.py:5-83)
# --------------------------------------------------------------------------------------------
class MelFilterBank:
    def __init__(self, spec_size, num_coefficients, sample_rate):
        nb = int(num_coefficients)
        max_mel = 2595.0 * math.log10(1.0 + (sample_rate / 2.0) / 700.0)
        mel_step = (max_mel - 0) / (nb + 1)
        edges = np.arange(0, nb + 2) * mel_step
        centers = [int(math.floor((math.floor(700.0 * (math.pow(10.0, m / 2595.0) - 1.0)) / (sample_rate / 2.0)) * spec_size))
                   for m in edges]
        fm = np.zeros((nb, spec_size))
        for i in range(nb):
            start, center, end = centers[i:i + 3]
            k1, k2 = float(center - start), float(end - center)
            with np.errstate(divide='ignore', invalid='ignore'):
                fm[i][start:center] = (np.array(range(start, center)) - start) / k1
                fm[i][center:end] = (end - np.array(range(center, end))) / k2
        self.melMatrix = fm.transpose()
        self.melMatrix = self._normal(self.melMatrix / self._norm_sum(self.melMatrix))
        self.melInvMatrix = self.melMatrix.transpose()
        self.melInvMatrix = self._normal(self.melInvMatrix / self._norm_sum(self.melInvMatrix))

    @staticmethod
    def _norm_sum(x):
        s = np.sum(x, axis=0)
        s[np.where(s == 0)] = 1.0
        return s

    @staticmethod
    def _normal(x):
        x[np.isnan(x)] = 0
        x[np.isinf(x)] = 0
        return x

    def toLogMels(self, spectrogram):
        return self._normal(np.log(np.dot(spectrogram, self.melMatrix) + 0.0000001))

    def fromLogMels(self, mel_spectrogram):
        return self._normal(np.dot(np.exp(mel_spectrogram), self.melInvMatrix))

    def fromMels(self, mel_spectrogram):
        return np.dot(mel_spectrogram, self.melInvMatrix)                 # MelFilterBank.py:60-61,75-76: no clean-up


# --------------------------------------------------------------------------------------------
# a8/a9: LDA decode + dequantise (R2)   LDASynthesis.py:19-28 ; Dequantization.py:15-18 ;
#        quantization.py:125-135
# --------------------------------------------------------------------------------------------
def pack_estimators(estimators, n_classes=9):
    """coef_/intercept_/classes_ of each fitted sklearn LDA -> dense W[b,k,F], bias[b,k], cls[b,k], n[b].
    Binary estimators carry a single score row (sklearn decision_function special case)."""
    nb = len(estimators)
    n_feat = estimators[0].coef_.shape[1]
    W = np.zeros((nb, n_classes, n_feat))
    b = np.full((nb, n_classes), -np.inf)
    cls = np.zeros((nb, n_classes))
    cnt = np.zeros(nb, dtype=np.int64)
    for i, est in enumerate(estimators):
        k = len(est.classes_)
        cnt[i] = k
        cls[i, :k] = est.classes_
        if k == 2:
            # scores = [-s, s]; argmax picks class 1 iff s > 0
            W[i, 1], b[i, 1] = est.coef_[0], est.intercept_[0]
            W[i, 0], b[i, 0] = 0.0, 0.0
        else:
            W[i, :k], b[i, :k] = est.coef_, est.intercept_
    return W, b, cls, cnt


def lda_predict(features, estimators, select):
    """Per-bin sklearn predict exactly as the node calls it (LDASynthesis.py:25-26), batched over frames."""
    x = np.asarray(features)[:, select]
    out = np.empty((x.shape[0], len(estimators)))
    for i, est in enumerate(estimators):
        out[:, i] = est.predict(x)
    return out


def lda_predict_packed(features, W, b, cls, select):
    """Closed form R2: label = cls[argmax_k (x[select] . W[b,k] + bias[b,k])] (first maximum wins)."""
    x = np.asarray(features, dtype=np.float64)[:, select]
    scores = np.einsum('tf,bkf->tbk', x, W) + b[None]
    idx = np.argmax(scores, axis=2)
    return np.take_along_axis(np.broadcast_to(cls[None], scores.shape), idx[..., None], axis=2)[..., 0], scores


def dequantize_spectrogram(q_spectrogram, medians_array):
    """quantization.py:125-135"""
    q = np.asarray(q_spectrogram).astype(int)
    out = np.zeros((q.shape[0], medians_array.shape[0]))
    for mel_bin in range(out.shape[1]):
        out[:, mel_bin] = medians_array[mel_bin][q[:, mel_bin]]
    return out


def dequantization_node(labels, medians_array):
    """Dequantization.py:15-18, one row per frame: lookup then gaussian_filter(sigma=0.5) across bins."""
    labels = np.atleast_2d(labels)
    c = np.arange(len(medians_array))
    out = np.empty((labels.shape[0], len(medians_array)))
    for t in range(labels.shape[0]):
        out[t] = gaussian_filter(medians_array[c, labels[t].astype(int)], sigma=0.5)
    return out


# --------------------------------------------------------------------------------------------
# a11-a13: streaming Griffin-Lim node (R3)   livenodes/GriffinLim.py:13-174
# --------------------------------------------------------------------------------------------
class GriffinLimNode:
    """Closed form of GriffinLimSynthesis for frameSize 16 ms / shift 10 ms style configurations.
    `noise[k]` is the np.random.rand(block_samples) draw of frame k (frames k >= spec_frames-1 draw)."""

    def __init__(self, frame_size_ms=16, frame_shift_ms=10, sample_rate=16000, n_mels=40, iterations=8,
                 cutoff=7900, norm_factor=1.0, use_log_mels=True):
        self.use_log_mels = use_log_mels
        fs, sh, sr = float(frame_size_ms), float(frame_shift_ms), float(sample_rate)
        self.sample_rate, self.frame_shift_ms = sr, sh
        self.fft_size = int((fs / 1000.0) * sr)
        self.hop = int((sh / 1000.0) * sr)
        self.context_width = int(fs / sh)
        self.block_len = self.context_width * 2 + 1
        self.block_samples = self.block_len * self.hop
        self.first_frame = self.block_len - self.context_width - 1      # GriffinLim.py:131 (0-based index)
        self.iterations = iterations
        self.norm = norm_factor
        self.window = np.blackman(self.fft_size)
        order = int((sr / 1000.0) * sh / 32.0)
        self.b, self.a = scipy.signal.iirfilter(order, float(cutoff) / float(sr / 2), btype="lowpass")
        self.mel = MelFilterBank(int(self.fft_size / 2 + 1), n_mels, sr)
        self.offsets = list(range(0, self.block_samples - self.fft_size, self.hop))   # Q3

    def block(self, logmel_frames, noise):
        """reconstructWavFromSpectrogram (GriffinLim.py:76-96) incl. quirk Q1 (no 1j in the phase term)."""
        spec = self.mel.fromLogMels(logmel_frames) if self.use_log_mels else self.mel.fromMels(logmel_frames)   # GriffinLim.py:84-87
        x = np.array(noise, dtype=np.float64, copy=True)
        n = self.fft_size
        for _ in range(self.iterations):
            X = np.array([np.fft.rfft(self.window * x[o:o + n]) for o in self.offsets])
            z = spec * np.exp(np.angle(X))
            re = np.zeros(self.block_samples)
            for j, o in enumerate(self.offsets):
                re[o:o + n] += np.real(np.fft.irfft(z[j])) * self.window
            x[:len(re)] = re
        return x

    def positions(self, n_frames):
        """Write-head position after each frame (GriffinLim.py:115-120).  int() truncation of
        (ms/1000)*sampleRate makes some hops 159 or 161 samples long (quirk Q7: first at frame 201)."""
        pos, ms = [], 0.0
        for _ in range(n_frames):
            ms += self.frame_shift_ms
            pos.append(int((ms / 1000.0) * self.sample_rate))
        return np.asarray(pos, dtype=np.int64)

    def synthesize(self, logmels, noise):
        """logmels (T x n_mels), noise (T x block_samples; row k used by frame k). Returns int16 PCM and
        the un-quantised low-passed signal.  The ring buffers of the node (GriffinLim.py:145-166) are
        restated on a linear time axis: a cell is only ever summed over the blocks that cover it after
        the write head passed it, so the ring never aliases (its length is > block + 2 hops)."""
        T = logmels.shape[0]
        nf = self.block_len - self.context_width                        # frames per block (2)
        bs = self.block_samples
        P = self.positions(T)
        wb = np.blackman(bs)
        buf = np.zeros(bs + (P[-1] if T else 0) + 1)                    # index = absolute position + bs
        win = np.zeros_like(buf)
        zi = scipy.signal.lfiltic(self.b, self.a, np.array([]))
        pcm, flt = [], []
        for k in range(self.first_frame, T):
            prev, pos = (P[k - 1] if k > 0 else 0), P[k]
            shifted = pos - prev
            blk = self.block(logmels[k - nf + 1:k + 1], noise[k])
            buf[pos:pos + bs] += blk                                    # [pos - bs, pos) in absolute terms
            win[pos:pos + bs] += wb
            num = buf[pos:pos + shifted].copy()
            den = win[pos:pos + shifted]
            nz = den != 0
            num[nz] = num[nz] / den[nz]
            y, zi = scipy.signal.lfilter(self.b, self.a, num, zi=zi)
            flt.append(y)
            pcm.append(np.int16(np.clip(y / (self.norm * 1.01), -0.99, 0.99) * (2 ** 15 - 1)))
        if not pcm:
            return np.zeros(0, dtype=np.int16), np.zeros(0)
        return np.hstack(pcm), np.hstack(flt)


# --------------------------------------------------------------------------------------------
# a14: batch Griffin-Lim  (local/offline.py:131-192), closed form R4
# --------------------------------------------------------------------------------------------
def griffin_lim_offline(spectrogram, noise, win_length=0.05, hop_size=0.01, num_iterations=8, return_float=False):
    """`noise` replaces np.random.rand(2*T*n_bins) (offline.py:164); only noise[:hop*(T-1)+win] matters."""
    audiosr = 16000
    win_len = int(win_length * audiosr)
    overlap = win_length / hop_size
    hop = int(win_len / overlap)
    n_bins = int(win_len / 2 + 1)
    mfb = MelFilterBank(n_bins, spectrogram.shape[1], 16000)
    spec = mfb.fromLogMels(spectrogram)
    T = spec.shape[0]
    w = np.hanning(win_len + 1)[:-1]
    x = np.array(noise, dtype=np.float64, copy=True)
    out_len = T * hop
    for _ in range(num_iterations):
        X = np.array([np.fft.rfft(w * x[i:i + win_len]) for i in range(0, hop * T, hop)])     # first T frames only
        z = spec * np.exp(1j * np.angle(X))
        re = np.zeros(out_len)
        for n, i in enumerate(range(0, out_len - win_len, hop)):
            re[i:i + win_len] += np.real(np.fft.irfft(z[n])) * w
        x[:out_len] = re
    rec = x[:out_len]
    scaled = np.int16(rec / np.max(np.abs(rec)) * 32767)
    return (scaled, rec) if return_float else scaled


# --------------------------------------------------------------------------------------------
# a15: audio -> log-mel target  (local/offline.py:219-241)
# --------------------------------------------------------------------------------------------
def compute_spectrogram(audio, sr=16000, window_length=0.05, window_shift=0.01, mel_bins=40):
    wl = int(sr * window_length)
    ws = int(sr * window_shift)
    overlap = wl - ws
    audio = np.hstack([np.zeros(overlap), audio])
    num_windows = int(np.floor((len(audio) - overlap) / ws))
    win = hann(wl)
    spec = np.zeros((num_windows, wl // 2 + 1))
    for i in range(num_windows):
        spec[i] = np.abs(np.fft.rfft(audio[i * ws:i * ws + wl] * win))
    return MelFilterBank(spec.shape[1], mel_bins, sr).toLogMels(spec).astype('float')


# --------------------------------------------------------------------------------------------
# a16: logistic quantisation  (local/quantization.py:83-135 ; train.py:78-93)
# --------------------------------------------------------------------------------------------
def compute_borders_logistic(spectrogram, nb_intervals):
    vmins = np.min(spectrogram, axis=0)
    vmaxs = np.max(spectrogram, axis=0)

    def sigmoid(t, vmin, vmax, k=0.5):
        L = abs(vmin) + vmax
        return L / (1 + np.exp(-k * t)) - abs(vmin)

    borders = np.zeros((spectrogram.shape[1], nb_intervals))
    medians = np.zeros((spectrogram.shape[1], nb_intervals))
    for b in range(spectrogram.shape[1]):
        y = sigmoid(np.linspace(-10, 10, nb_intervals + 1, endpoint=True), vmins[b], vmaxs[b])
        borders[b, :-1] = y[1:-1]
        borders[b, -1] = vmaxs[b]
        medians[b, :] = sigmoid(np.linspace(-9.5, 9.5, nb_intervals, endpoint=True), vmins[b], vmaxs[b])
    return medians, borders


def quantize_spectrogram(spectrogram, borders):
    q = np.zeros(spectrogram.shape)
    for mel_bin in range(spectrogram.shape[1]):
        for interval_nb in reversed(range(borders.shape[1])):
            q[np.where(spectrogram[:, mel_bin] <= borders[mel_bin, interval_nb]), mel_bin] = interval_nb
    return q


# --------------------------------------------------------------------------------------------
# a17-a19: training  (train.py:96-168)
# --------------------------------------------------------------------------------------------
def feature_selection(x_train, y_train, nb_feats=150):
    from scipy.stats import spearmanr
    cs = np.zeros(x_train.shape[1])
    target = np.mean(y_train, axis=1)
    for f in range(x_train.shape[1]):
        if np.isclose(np.sum(x_train[:, f]), 0):
            cs[f] = 0
            continue
        cs[f], _ = spearmanr(x_train[:, f], target)
    return np.argsort(np.abs(cs))[np.max([-nb_feats, -len(cs)]):], cs


def train(eeg, audio16k, sfreq_eeg, bad_channels, nb_mel_bins=40, nb_intervals=9, nb_feats=150):
    """train.train (train.py:132-168) from already-decimated 16 kHz audio (the decimate call at
    train.py:125 is audio-side preparation)."""
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    eeg = np.asarray(eeg, dtype=np.float64)
    if len(bad_channels) > 0:
        mask = np.ones(eeg.shape[1], bool)
        mask[bad_channels] = False
        eeg = eeg[:, mask]
    x_train = herff2016_b(eeg, sfreq_eeg, 0.05, 0.01)
    y_train = compute_spectrogram(audio16k, 16000, 0.016, 0.01, nb_mel_bins)
    y_train = y_train[20:-4]                                             # train.py:144
    medians, borders = compute_borders_logistic(y_train, nb_intervals)
    q = quantize_spectrogram(y_train, borders)
    select, _ = feature_selection(x_train, y_train, nb_feats)
    x_train = x_train[:, select]
    n = min(len(x_train), len(q))
    x_train, q = x_train[:n], q[:n]
    estimators = [LinearDiscriminantAnalysis() for _ in range(nb_mel_bins)]
    for b in range(nb_mel_bins):
        estimators[b].fit(x_train, q[:, b])
    return x_train, q, medians, estimators, select


# --------------------------------------------------------------------------------------------
# function-level batch decode = CPU baseline (ii) of SURVEY.md 8d
# --------------------------------------------------------------------------------------------
def decode_offline_batch(eeg, sr, estimators, select, medians, noise, gl_iterations=8):
    x = herff2016_b(eeg, sr)
    labels = lda_predict(x, estimators, select)
    spec = dequantize_spectrogram(labels, medians)
    return labels, spec, griffin_lim_offline(spec, noise, num_iterations=gl_iterations)


def decode_streaming(eeg, sr, estimators, select, medians, noise, gl_norm=10, chunk_size=32, gl_iterations=8):
    """Closed form of decode.setup_decoder's chain (decode.py:152-168) after all data was pushed."""
    x = ecog_feat_calc(eeg, sr, 50, 10, 4, 5, 50, chunk_size)
    labels = lda_predict(x, estimators, select)
    spec = dequantization_node(labels, medians)
    gl = GriffinLimNode(16, 10, 16000, spec.shape[1], gl_iterations, norm_factor=gl_norm)
    pcm, flt = gl.synthesize(spec, noise)
    return x, labels, spec, pcm, flt
