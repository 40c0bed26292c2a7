"""Training-side operators (train.py:78-168 of the reference) on top of libsgs.

Data passes (min/max, quantisation, ranking, Gram matrix, class sums) run on the device; what is left on the
host is O(bins x features^3) dense algebra on 150 x 150 matrices (the eigen-solves sklearn's SVD solver hides)
and the construction of estimator objects.  With torch.distributed initialised the row-sharded statistics are
summed with one all-reduce (NCCL on GPUs; gloo in the CPU tests) before the solve."""
import numpy as np

from . import _lib

TOL = 1e-4                      # sklearn LinearDiscriminantAnalysis default


def _dev(a, dtype='float64'):
    """numpy -> contiguous host array; torch CUDA tensor -> contiguous tensor of the wanted dtype (stays on the device)."""
    if _lib._is_torch(a):
        import torch
        return a.to(getattr(torch, dtype)).contiguous()
    return np.ascontiguousarray(a, dtype=getattr(np, dtype))


def _empty_like_kind(ref, shape, dtype='float64'):
    if _lib._is_torch(ref):
        import torch
        return torch.empty(shape, dtype=getattr(torch, dtype), device=ref.device)
    return np.empty(shape, dtype=getattr(np, dtype))


def _to_host(a):
    return a.cpu().numpy() if _lib._is_torch(a) else a


# ---- quantisation (train.py:78-93) ----------------------------------------------------------------------
def quantization(y_train, nb_intervals=9):
    """y_train (N x bins), numpy or torch-CUDA.  medians / borders are small host arrays; the labels stay where y lives."""
    from local.quantization import compute_borders_logistic
    y = _dev(y_train)
    _lib.ensure_init()
    mn, mx = _empty_like_kind(y, (y.shape[1],)), _empty_like_kind(y, (y.shape[1],))
    st = _lib.current_stream(y)
    _lib.check(_lib.lib().sgs_col_minmax(_lib.ptr(y), y.shape[0], y.shape[1], _lib.ptr(mn), _lib.ptr(mx), st))
    medians, borders = compute_borders_logistic(np.vstack([_to_host(mn), _to_host(mx)]), nb_intervals)   # only min/max enter the formula
    borders = np.ascontiguousarray(borders)
    q = _empty_like_kind(y, tuple(y.shape))
    _lib.check(_lib.lib().sgs_quantize(_lib.ptr(y), y.shape[0], y.shape[1], _lib.ptr(borders), nb_intervals, _lib.ptr(q), st))
    return medians, borders, q


# ---- feature selection (train.py:96-109) ------------------------------------------------------------------
def spearman(x_train, y_train):
    x, y = _dev(x_train), _dev(y_train)
    n = min(len(x), len(y))
    _lib.ensure_init()
    rho, colsum = np.empty(x.shape[1]), np.empty(x.shape[1])
    _lib.check(_lib.lib().sgs_spearman(_lib.ptr(x), n, x.shape[1], x.shape[1], _lib.ptr(y), y.shape[1], _lib.ptr(rho),
                                       _lib.ptr(colsum), _lib.current_stream(x)))
    return rho, colsum


def feature_selection(x_train, y_train, nb_feats=150):
    if len(x_train) != len(y_train):
        # scipy.stats.spearmanr raises on unequal lengths; keep the reference's failure mode
        raise ValueError("all the input array dimensions must match: %d feature rows vs %d target rows" % (len(x_train), len(y_train)))
    cs, colsum = spearman(x_train, y_train)
    cs[np.isclose(colsum, 0)] = 0
    return np.argsort(np.abs(cs))[np.max([-nb_feats, -len(cs)]):]


# ---- LDA fit from sufficient statistics (train.py:112-118, closed form R5) --------------------------------
def lda_stats(x_train, select, labels, n_classes=9, xbar=None):
    """x_train (N x width) full stacked features, select -> model columns, labels (N x bins); numpy or torch-CUDA.
    The statistics come back as (small) host arrays."""
    x, lab = _dev(x_train), _dev(labels)
    sel = np.ascontiguousarray(select, dtype=np.int32)
    n, nf, nb = len(x), len(sel), lab.shape[1]
    _lib.ensure_init()
    out_xbar = np.empty(nf); G = np.empty((nf, nf)); sums = np.empty((nb, n_classes, nf)); counts = np.empty((nb, n_classes))
    xin = None if xbar is None else np.ascontiguousarray(xbar, dtype=np.float64)
    _lib.check(_lib.lib().sgs_lda_stats(_lib.ptr(x), n, x.shape[1], _lib.ptr(sel), nf, _lib.ptr(lab), nb, n_classes, _lib.ptr(xin),
                                        _lib.ptr(out_xbar), _lib.ptr(G), _lib.ptr(sums), _lib.ptr(counts), _lib.current_stream(x)))
    return dict(n=float(n), xbar=out_xbar, G=G, sums=sums, counts=counts)


def col_means(x_train, select):
    x = _dev(x_train)
    sel = np.ascontiguousarray(select, dtype=np.int32)
    _lib.ensure_init()
    out = np.empty(len(sel))
    _lib.check(_lib.lib().sgs_col_means(_lib.ptr(x), len(x), x.shape[1], _lib.ptr(sel), len(sel), _lib.ptr(out), _lib.current_stream(x)))
    return out


def allreduce_stats(stats, group=None):
    """Sum row-sharded statistics over the process group (no-op without torch.distributed).  All shards must have
    been centred on the same xbar (see distributed_lda_stats)."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return stats
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    flat = np.concatenate([[stats['n']], stats['G'].ravel(), stats['sums'].ravel(), stats['counts'].ravel()])
    t = torch.from_numpy(flat).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    flat = t.cpu().numpy()
    nf = stats['G'].shape[0]
    o = 1
    out = dict(stats)
    out['n'] = float(flat[0])
    out['G'] = flat[o:o + nf * nf].reshape(nf, nf); o += nf * nf
    out['sums'] = flat[o:o + stats['sums'].size].reshape(stats['sums'].shape); o += stats['sums'].size
    out['counts'] = flat[o:o + stats['counts'].size].reshape(stats['counts'].shape)
    return out


def global_mean(local_mean, n_local, group=None):
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return local_mean
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_mean
    dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    t = torch.from_numpy(np.concatenate([[float(n_local)], local_mean * n_local])).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu().numpy()
    return t[1:] / t[0]


def distributed_lda_stats(x_local, select, labels_local, n_classes=9, group=None):
    """Row shard -> globally summed statistics: all-reduce the mean, centre every shard on it, all-reduce the rest."""
    xbar = global_mean(col_means(x_local, select), len(x_local), group)
    return allreduce_stats(lda_stats(x_local, select, labels_local, n_classes, xbar), group)


class PackedLDA:
    """Minimal estimator (coef_/intercept_/classes_/predict) used when scikit-learn is not importable."""

    def __init__(self):
        self.coef_ = self.intercept_ = self.classes_ = None

    def decision_function(self, X):
        s = np.asarray(X) @ self.coef_.T + self.intercept_
        return s.ravel() if s.shape[1] == 1 else s

    def predict(self, X):
        s = self.decision_function(X)
        idx = (s > 0).astype(int) if s.ndim == 1 else s.argmax(axis=1)
        return self.classes_[idx]


def _new_estimator():
    try:
        from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
        return LinearDiscriminantAnalysis()
    except ImportError:
        return PackedLDA()


def fit_from_stats(stats, tol=TOL):
    """sklearn's _solve_svd (discriminant_analysis.py) restated on (G, class sums, counts): the SVD of the scaled,
    within-class-centred data matrix is replaced by the eigen-decomposition of its 150 x 150 Gram matrix."""
    N, xbar, G = stats['n'], stats['xbar'], stats['G']
    estimators = []
    for b in range(stats['sums'].shape[0]):
        present = np.nonzero(stats['counts'][b] > 0)[0]
        n_k = stats['counts'][b][present]
        K = len(present)
        priors = n_k / N
        d = stats['sums'][b][present] / n_k[:, None]                   # class means minus the global mean
        means = xbar + d
        xbar_b = priors @ means
        SW = G - (d * n_k[:, None]).T @ d
        std = np.sqrt(np.maximum(np.diag(SW), 0.0) / N)
        std[std == 0] = 1.0
        fac = 1.0 / (N - K)
        C = (SW / np.outer(std, std)) * fac
        evals, evecs = np.linalg.eigh((C + C.T) * 0.5)
        order = np.argsort(evals)[::-1]
        S = np.sqrt(np.maximum(evals[order], 0.0))
        V = evecs[:, order]
        rank = int(np.sum(S > tol))
        scalings = (V[:, :rank] / std[:, None]) / S[:rank]
        fac2 = 1.0 if K == 1 else 1.0 / (K - 1)
        X2 = (np.sqrt((N * priors) * fac2)[:, None] * (means - xbar_b)) @ scalings
        _, S2, Vt2 = np.linalg.svd(X2, full_matrices=False)
        rank2 = int(np.sum(S2 > tol * S2[0]))
        scalings_ = scalings @ Vt2.T[:, :rank2]
        coef = (means - xbar_b) @ scalings_
        intercept = -0.5 * np.sum(coef ** 2, axis=1) + np.log(priors)
        coef_ = coef @ scalings_.T
        intercept_ = intercept - xbar_b @ coef_.T
        est = _new_estimator()
        est.classes_ = present.astype(np.float64)
        est.priors_ = priors
        est.means_ = means
        est.xbar_ = xbar_b
        est.scalings_ = scalings_
        est._max_components = min(K - 1, G.shape[0])
        est.explained_variance_ratio_ = (S2 ** 2 / np.sum(S2 ** 2))[:est._max_components]
        est.n_features_in_ = G.shape[0]
        if K == 2:
            est.coef_ = np.array(coef_[1, :] - coef_[0, :], ndmin=2)
            est.intercept_ = np.array(intercept_[1] - intercept_[0], ndmin=1)
        else:
            est.coef_, est.intercept_ = coef_, intercept_
        estimators.append(est)
    return estimators
