"""Audio -> log-mel spectrogram on the device (sgs_logmel), the target side of train.py (local/offline.py:219-241)."""
import numpy as np
from scipy.signal.windows import hann

from . import _lib
from .design import MelTables


def log_mel_spectrogram(audio, sr=16000, window_length=0.05, window_shift=0.01, mel_bins=40):
    """audio: 1-D numpy array or torch-CUDA tensor; the (frames x mel_bins) result lives where the audio does."""
    win_len = int(sr * window_length)
    shift = int(sr * window_shift)
    overlap = win_len - shift
    is_torch = _lib._is_torch(audio)
    if is_torch:
        import torch
        audio = audio.to(torch.float64).contiguous()
    else:
        audio = np.ascontiguousarray(audio, dtype=np.float64)
    n_frames = int(np.floor((len(audio) + overlap - overlap) / shift))
    mel = MelTables(win_len // 2 + 1, mel_bins, sr)
    window = np.ascontiguousarray(hann(win_len), dtype=np.float64)
    m = np.ascontiguousarray(mel.mel, dtype=np.float64)
    if is_torch:
        out = torch.empty((n_frames, mel_bins), dtype=torch.float64, device=audio.device)
    else:
        out = np.empty((n_frames, mel_bins), dtype=np.float64)
    _lib.ensure_init()
    _lib.check(_lib.lib().sgs_logmel(_lib.ptr(audio), len(audio), _lib.ptr(window), win_len, shift, _lib.ptr(m), m.shape[0],
                                     mel_bins, n_frames, _lib.ptr(out), _lib.current_stream(audio)))
    return out


def decimate(audio, q, n=8):
    """scipy.signal.decimate(audio, q) (IIR, zero phase) on the device: train.py:125 brings 48 kHz audio to 16 kHz with it.
    Follows scipy's own construction: cheby1(n, 0.05, 0.8/q) as second-order sections, sosfiltfilt with its odd
    extension and steady-state initial conditions, then every q-th sample."""
    from scipy.signal import cheby1, sosfilt_zi
    is_torch = _lib._is_torch(audio)
    if is_torch:
        import torch
        audio = audio.to(torch.float64).contiguous()
    else:
        audio = np.ascontiguousarray(audio, dtype=np.float64)
    if audio.ndim != 1:
        raise ValueError("decimate expects a 1-D signal")
    q = int(q)
    sos = np.ascontiguousarray(cheby1(n, 0.05, 0.8 / q, output='sos'), dtype=np.float64)
    n_sections = sos.shape[0]
    ntaps = 2 * n_sections + 1
    ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    edge = 3 * ntaps
    if len(audio) <= edge:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % edge)
    zi = np.ascontiguousarray(sosfilt_zi(sos), dtype=np.float64)
    radius = max(abs(np.roots(sec[3:])).max() for sec in sos)
    warm = int(np.ceil(70.0 * np.log(2.0) / -np.log(radius))) + 64
    n_out = (len(audio) + q - 1) // q
    out = torch.empty(n_out, dtype=torch.float64, device=audio.device) if is_torch else np.empty(n_out, dtype=np.float64)
    _lib.ensure_init()
    _lib.check(_lib.lib().sgs_decimate(_lib.ptr(audio), len(audio), q, _lib.ptr(sos), _lib.ptr(zi), n_sections, edge, warm,
                                       _lib.ptr(out), _lib.current_stream(audio)))
    return out
