"""Logistic quantisation of the log-mel target (reference: local/quantization.py:83-135).

compute_borders_logistic is 40 x 9 scalar evaluations (host).  quantize_spectrogram / dequantize_spectrogram
act on (frames x bins) arrays; dequantisation runs on the device (sgs_dequantize)."""
import numpy as np

from sgs import _lib


def compute_borders_logistic(spectrogram, nb_intervals):
    vmins = np.min(spectrogram, axis=0)
    vmaxs = np.max(spectrogram, axis=0)

    def sigmoid(t, vmin, vmax, k=0.5):
        L = abs(vmin) + vmax
        return L / (1 + np.exp(-k * t)) - abs(vmin)

    nb = spectrogram.shape[1]
    borders_array = np.zeros((nb, nb_intervals))
    medians_array = np.zeros((nb, nb_intervals))
    t_b = np.linspace(-10, 10, nb_intervals + 1, endpoint=True)
    t_m = np.linspace(-9.5, 9.5, nb_intervals, endpoint=True)
    for b in range(nb):
        y = sigmoid(t_b, vmins[b], vmaxs[b])
        borders_array[b, :-1] = y[1:-1]
        borders_array[b, -1] = vmaxs[b]
        medians_array[b, :] = sigmoid(t_m, vmins[b], vmaxs[b])
    return medians_array, borders_array


def quantize_spectrogram(spectrogram, borders):
    """Label = smallest interval whose border is >= the value (0 when above every border), as the reference's
    descending overwrite loop yields (quantization.py:112-122)."""
    spectrogram = np.asarray(spectrogram)
    q = np.zeros(spectrogram.shape)
    for b in range(spectrogram.shape[1]):
        col = spectrogram[:, b]
        for k in reversed(range(borders.shape[1])):
            q[col <= borders[b, k], b] = k
    return q


def dequantize_spectrogram(q_spectrogram, medians_array):
    labels = np.ascontiguousarray(np.asarray(q_spectrogram).astype(int).astype(np.float64))
    med = np.ascontiguousarray(medians_array, dtype=np.float64)
    out = np.empty((labels.shape[0], med.shape[0]), dtype=np.float64)
    if labels.shape[0]:
        _lib.ensure_init()
        _lib.check(_lib.lib().sgs_dequantize(_lib.ptr(med), med.shape[0], med.shape[1], None, 0, _lib.ptr(labels),
                                             labels.shape[0], 0, _lib.ptr(out), None))
    return out
