"""LDA decoding + dequantisation operator (sgs_lda_decode) and estimator packing.

An estimator is anything exposing coef_, intercept_ and classes_ the way a fitted
sklearn.discriminant_analysis.LinearDiscriminantAnalysis does (livenodes/LDASynthesis.py:15 unpickles a list of
them); the weights are packed once into dense device arrays."""
import numpy as np

from . import _lib
from .design import gaussian_taps

N_CLASSES = 9       # train.py:150 quantises into 9 intervals; the kernel is built for that width


def pack_estimators(estimators, n_classes=N_CLASSES):
    """-> W[b,k,F], bias[b,k] (-inf where a bin lacks a class), cls[b,k].
    sklearn's binary special case (one score row, predict = score > 0) becomes the two rows {0, coef_}."""
    nb = len(estimators)
    n_feat = int(np.asarray(estimators[0].coef_).shape[1])
    W = np.zeros((nb, n_classes, n_feat))
    b = np.full((nb, n_classes), -np.inf)
    cls = np.zeros((nb, n_classes))
    for i, est in enumerate(estimators):
        classes = np.asarray(est.classes_, dtype=np.float64)
        coef = np.asarray(est.coef_, dtype=np.float64)
        icpt = np.asarray(est.intercept_, dtype=np.float64)
        k = len(classes)
        if k > n_classes:
            raise ValueError("estimator %d has %d classes; at most %d are supported" % (i, k, n_classes))
        cls[i, :k] = classes
        if k == 2 and coef.shape[0] == 1:
            W[i, 1], b[i, 1] = coef[0], icpt[0]
            W[i, 0], b[i, 0] = 0.0, 0.0
        else:
            W[i, :k], b[i, :k] = coef, icpt
    return W, b, cls


class LdaDecoder:
    """40 per-bin classifiers + medians table bound to a device model."""

    def __init__(self, estimators, select, medians, sigma=0.5):
        if isinstance(estimators, tuple):
            self.W, self.bias, self.cls = (np.ascontiguousarray(a, dtype=np.float64) for a in estimators)
        else:
            self.W, self.bias, self.cls = pack_estimators(estimators)
        self.select = np.ascontiguousarray(select, dtype=np.int32)
        self.medians = np.ascontiguousarray(medians, dtype=np.float64)
        self.taps = np.ascontiguousarray(gaussian_taps(sigma), dtype=np.float64)
        if self.W.shape[2] != len(self.select):
            raise ValueError("model has %d features but select has %d entries" % (self.W.shape[2], len(self.select)))
        if self.medians.shape[0] != self.W.shape[0]:
            raise ValueError("medians_array rows (%d) != number of estimators (%d)" % (self.medians.shape[0], self.W.shape[0]))
        self.n_bins = self.W.shape[0]
        self._handle = None

    def handle(self):
        if self._handle is None:
            _lib.ensure_init()
            h = _lib.c_void_p()
            _lib.check(_lib.lib().sgs_lda_model_create(
                _lib.C.byref(h), self.n_bins, self.W.shape[1], self.W.shape[2], _lib.ptr(self.W), _lib.ptr(self.bias),
                _lib.ptr(self.cls), _lib.ptr(self.select), _lib.ptr(self.medians), self.medians.shape[1],
                _lib.ptr(self.taps), len(self.taps) // 2))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().sgs_lda_model_destroy(self._handle)
        except Exception:
            pass

    def last_rescored(self):
        """(frame, bin) pairs the last tensor-core decode re-scored exactly in fp64 (0 when the fp64 kernel ran alone)."""
        n = _lib.c_int(0)
        _lib.check(_lib.lib().sgs_lda_last_rescored(self.handle(), _lib.C.byref(n)))
        return n.value

    def decode(self, feat, order=0, step=1, first_row=0, n_rows=None, smooth=False, want_labels=True, want_spec=True):
        """feat: un-stacked log-power (.., W, C) with (order, step, first_row) describing the stacked view, or
        already stacked rows (.., R, 5C) with order=0.  Returns (labels, spectrogram), each (.., rows, n_bins)."""
        is_torch = _lib._is_torch(feat)
        squeeze = feat.ndim == 2
        if squeeze:
            feat = feat[None]
        S, nw, Cn = feat.shape
        rows = (nw - first_row) if n_rows is None else int(n_rows)
        rows = max(rows, 0)
        if is_torch:
            import torch
            feat = feat.contiguous()
            assert feat.dtype == torch.float64
            mk = lambda: torch.empty((S, rows, self.n_bins), dtype=torch.float64, device=feat.device)
        else:
            feat = np.ascontiguousarray(feat, dtype=np.float64)
            mk = lambda: np.empty((S, rows, self.n_bins), dtype=np.float64)
        labels = mk() if want_labels else None
        spec = mk() if want_spec else None
        if rows > 0:
            _lib.check(_lib.lib().sgs_lda_decode(self.handle(), _lib.ptr(feat), S, nw, Cn, rows, first_row, order, step,
                                                 _lib.ptr(labels), _lib.ptr(spec), int(bool(smooth)), _lib.current_stream(feat)))
        if squeeze:
            labels = labels[0] if labels is not None else None
            spec = spec[0] if spec is not None else None
        return labels, spec
