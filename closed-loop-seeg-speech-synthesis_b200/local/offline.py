"""Batch numeric entry points with the reference's signatures (local/offline.py), backed by libsgs.

    herff2016_b          local/offline.py:12-128   high-gamma log-power features (+ temporal stacking)
    griffin_lim          local/offline.py:131-192  batch Griffin-Lim (800-point frames, complex phase)
    pearson_correlation  local/offline.py:195-216  evaluation metric: per-bin Pearson r between two spectrograms
    compute_spectrogram  local/offline.py:219-241  audio -> 40-bin log-mel target

Arrays go in and come out as numpy (host) or as torch CUDA tensors (resident); nothing is computed on the CPU
except the O(1)-per-configuration tables in sgs.design."""
import numpy as np

from sgs.features import FeatureExtractor

_extractors = {}


def _extractor(sr, window_length, window_shift, line_noise):
    key = (sr, window_length, window_shift, 50 if line_noise == 50 else 60)
    if key not in _extractors:
        _extractors[key] = FeatureExtractor(sr, window_length, window_shift, key[3])
    return _extractors[key]


def herff2016_b(eeg, sr, window_length=0.05, window_shift=0.01, line_noise=50, skip_stacking=False):
    """Offline computation of the Herff et al. 2016 feature paradigm, compatible with the warm start of the node
    based system.  eeg: samples x channels.  Returns windows x channels log-power, or (windows-20) x 5*channels
    with the 5-tap temporal context stacked (column c*5+tap, tap 0 oldest)."""
    fe = _extractor(sr, window_length, window_shift, line_noise)
    feat = fe.log_power(eeg)
    return feat if skip_stacking else fe.stack(feat)


def griffin_lim(spectrogram, win_length=0.05, hop_size=0.01, num_iterations=8, noise=None):
    """Reconstruct an audible acoustic signal using the Griffin-Lim approach (frames x mel bins -> int16 audio).
    `noise` replaces the reference's np.random.rand(2*T*n_bins) start; by default it is drawn from numpy's global
    stream exactly as the reference does."""
    from sgs.griffinlim import griffin_lim_batch
    spectrogram = np.asarray(spectrogram, dtype=np.float64)
    if noise is None:
        win_len = int(win_length * 16000)
        n_bins = int(win_len / 2 + 1)
        noise = np.random.rand(spectrogram.shape[0] * n_bins * 2)
    return griffin_lim_batch(spectrogram[None], np.asarray(noise)[None], win_length, hop_size, num_iterations)[0]


def compute_spectrogram(audio, sr=16000, window_length=0.05, window_shift=0.01, mel_bins=40):
    from sgs.spectrogram import log_mel_spectrogram
    return log_mel_spectrogram(audio, sr, window_length, window_shift, mel_bins)


def pearson_correlation(spectrogram_1, spectrogram_2, return_means=False):
    """Mean and standard deviation over the mel bins of the per-bin Pearson correlation between two spectrograms
    (frames x bins); with return_means also the list of per-bin values.  Evaluation metric, O(frames x bins)."""
    if isinstance(spectrogram_1, str):
        spectrogram_1 = np.load(spectrogram_1)
    if isinstance(spectrogram_2, str):
        spectrogram_2 = np.load(spectrogram_2)
    a, b = np.asarray(spectrogram_1, dtype=np.float64), np.asarray(spectrogram_2, dtype=np.float64)
    assert a.shape == b.shape, 'Shapes of spectrograms do not match.'
    a = a - a.mean(axis=0)
    b = b - b.mean(axis=0)
    with np.errstate(invalid='ignore', divide='ignore'):
        rs = list((a * b).sum(axis=0) / np.sqrt((a * a).sum(axis=0) * (b * b).sum(axis=0)))
    mean, std = np.mean(rs), np.std(rs)
    return (mean, std, rs) if return_means else (mean, std)
