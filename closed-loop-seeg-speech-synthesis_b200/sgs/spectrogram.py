"""Audio -> log-mel spectrogram on the device (sgs_logmel), the target side of train.py (local/offline.py:219-241)."""
import numpy as np
from scipy.signal.windows import hann

from . import _lib
from .design import MelTables


def log_mel_spectrogram(audio, sr=16000, window_length=0.05, window_shift=0.01, mel_bins=40):
    win_len = int(sr * window_length)
    shift = int(sr * window_shift)
    overlap = win_len - shift
    audio = np.ascontiguousarray(audio, dtype=np.float64)
    n_frames = int(np.floor((len(audio) + overlap - overlap) / shift))
    mel = MelTables(win_len // 2 + 1, mel_bins, sr)
    window = np.ascontiguousarray(hann(win_len), dtype=np.float64)
    m = np.ascontiguousarray(mel.mel, dtype=np.float64)
    out = np.empty((n_frames, mel_bins), dtype=np.float64)
    _lib.ensure_init()
    _lib.check(_lib.lib().sgs_logmel(_lib.ptr(audio), len(audio), _lib.ptr(window), win_len, shift, _lib.ptr(m), m.shape[0],
                                     mel_bins, n_frames, _lib.ptr(out), None))
    return out
