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


# ---- train.train as a multi-GPU job ---------------------------------------------------------------------
# One process per GPU under torch.distributed, every rank calling train.train with the same host arrays:
#   * features + Spearman are sharded by CHANNEL BLOCK (channels are independent through filtering and windowing, and a
#     rank correlation needs a column's whole time axis): rank r uploads and filters only its channels - the upload of the
#     recording, the largest single cost of a 1 h session, shrinks by the world size;
#   * the audio side (decimation, log-mel target, logistic borders, labels) is replicated: the Spearman target (the frame
#     mean of the log-mel spectrogram) is needed whole on every rank, so there is no min/max exchange to make;
#   * the per-column correlations are all-gathered (5 doubles per channel) and every rank derives the same `select`;
#   * the selected columns, spread over the ranks that own their channels, are summed into one (N x 150) matrix - each column
#     is non-zero on exactly one rank, so the sum is exact;
#   * the LDA statistics are sharded by ROW: mean all-reduce, then ONE all-reduce of (n, G, class sums, counts);
#   * the 40 eigen-problems are dealt round-robin and the fitted estimators all-gathered.
# The orchestration takes its array operators as an object so that the CPU tests can drive it with numpy restatements over
# gloo (tests/test_train_host.py); DeviceOps below is the product binding.

def _dist(group=None):
    if group is False:
        return None, 0, 1
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        pass
    return None, 0, 1


def block_shard(n, rank, world):
    """[lo, hi) of a contiguous balanced split of n items (sizes differ by at most one) - decode.session_shard's rule."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _allreduce_sum_(a, group=None):
    """In-place sum over the group of a numpy array or torch tensor; returns (bytes, seconds) of the collective."""
    import time
    dist, _, world = _dist(group)
    if world == 1:
        return 0, 0.0
    import torch
    t = a if _lib._is_torch(a) else torch.from_numpy(a)
    on_gpu = dist.get_backend(group) == 'nccl'
    buf = t.cuda() if on_gpu and not t.is_cuda else t
    if buf.is_cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if buf.is_cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if buf is not t:
        t.copy_(buf)
    return int(t.numel() * t.element_size()), dt


def _allgather_objects(obj, group=None):
    dist, _, world = _dist(group)
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


def _torch_dtype(dtype):
    import torch
    return torch.from_numpy(np.empty(0, dtype=dtype)).dtype


class DeviceOps:
    """The device operators train.train is made of (libsgs through local/offline.py, sgs/spectrogram.py and this module)."""

    def upload(self, a, dtype=None):
        if _lib._is_torch(a):                   # already resident (the many-fold driver keeps the recording on the device)
            import torch
            return (a if dtype is None else a.to(_torch_dtype(dtype))).contiguous()
        from . import hostio
        return hostio.upload(a, dtype)          # chunked through page-locked slots on several threads; gathers a strided channel block

    def sync(self):
        import torch
        torch.cuda.synchronize()

    def features(self, eeg, sfreq_eeg):
        from local.offline import herff2016_b
        return herff2016_b(eeg, sfreq_eeg, 0.05, 0.01)

    def target(self, audio, audio_sr):
        from local.offline import compute_spectrogram
        if audio_sr != 16000:
            from .spectrogram import decimate
            audio = decimate(audio, int(round(audio_sr / 16000)))
        return compute_spectrogram(audio, 16000, 0.016, 0.01)

    def quantization(self, y, nb_intervals):
        return quantization(y, nb_intervals)

    def spearman(self, x, y):
        return spearman(x, y)

    def zeros(self, shape, like):
        import torch
        return torch.zeros(shape, dtype=torch.float64, device=like.device)

    def allgather_1d(self, local, n, group=None):
        """The block_shard pieces of a length-n vector, one per rank, concatenated on every rank (NCCL all-gather of equal
        padded chunks, then re-packed: the pieces differ in length by at most one)."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group)
        chunk = -(-int(n) // world)
        buf = torch.zeros(chunk, dtype=local.dtype, device=local.device)
        buf[:len(local)] = local
        out = torch.empty(world * chunk, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, buf, group=group)
        sizes = [block_shard(n, r, world) for r in range(world)]
        if all(hi - lo == chunk for lo, hi in sizes):
            return out
        return torch.cat([out[r * chunk:r * chunk + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])

    def index(self, idx, like):
        import torch
        return torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(like.device)

    def contiguous(self, a):
        return a.contiguous()

    def to_host(self, a):
        from . import hostio
        return hostio.download(a)

    def col_means(self, x, select):
        return col_means(x, select)

    def lda_stats(self, x, select, labels, n_classes, xbar):
        return lda_stats(x, select, labels, n_classes, xbar)


last_profile = {}       # stage times / collective sizes of the last sharded_fit in this process (bench.py's config3 record)


def sharded_fit(eeg, audio, sfreq_eeg, sfreq_audio, nb_mel_bins=40, nb_intervals=9, nb_feats=150, ops=None, group=None,
                distributed=True):
    """train.py:132-168 after the bad-channel mask, on `world` ranks (see the block comment above; world = 1 is the
    single-GPU path).  eeg (T x C) and audio are HOST arrays, the same on every rank (device tensors are taken as they are).
    distributed=False fits on this rank alone even when a process group exists (the many-fold driver shards whole folds).
    Returns (x_train[:, select] (device / ops array), labels, medians, estimators, select)."""
    import time
    ops = ops or DeviceOps()
    dist, rank, world = _dist(group) if distributed else (None, 0, 1)
    if not distributed:
        group = False                       # the collectives below become no-ops
    prof = {'world': world}
    t_all = time.perf_counter()

    def lap(key, t0):
        ops.sync()
        prof[key] = prof.get(key, 0.0) + (time.perf_counter() - t0)

    C = eeg.shape[1]
    c0, c1 = block_shard(C, rank, world)
    t0 = time.perf_counter()
    eeg_d = ops.upload(eeg[:, c0:c1])
    if world > 1 and not _lib._is_torch(audio):
        # every rank needs the whole audio (the target spectrogram is computed replicated), but not over its own PCIe link and
        # out of the one host memory all ranks share: each uploads 1/world of it and the pieces are summed into place over NVLink
        a_lo, a_hi = block_shard(len(audio), rank, world)
        piece = ops.upload(np.asarray(audio)[a_lo:a_hi], np.float64)
        prof['h2d_bytes'] = int(eeg[:, c0:c1].nbytes + (a_hi - a_lo) * 8)
        lap('h2d_s', t0)
        t1 = time.perf_counter()
        if hasattr(ops, 'allgather_1d'):
            audio_d = ops.allgather_1d(piece, len(audio), group)            # all-gather: 1/world of the all-reduce's traffic
        else:
            audio_d = ops.zeros((len(audio),), like=eeg_d)
            audio_d[a_lo:a_hi] = piece
            _allreduce_sum_(audio_d, group)
        ops.sync()
        prof['audio_allreduce_bytes'], prof['audio_allreduce_s'] = int(len(audio) * 8), time.perf_counter() - t1
        del piece
        t0 = time.perf_counter()
    else:
        audio_d = ops.upload(audio, np.float64)
        prof['h2d_bytes'] = 0 if _lib._is_torch(eeg) else int(eeg[:, c0:c1].nbytes + np.asarray(audio).size * 8)
        prof['audio_allreduce_bytes'], prof['audio_allreduce_s'] = 0, 0.0
    lap('h2d_s', t0)

    t0 = time.perf_counter()
    x_r = ops.features(eeg_d, sfreq_eeg) if c1 > c0 else None          # (N x 5 (c1 - c0)), column (c - c0) * 5 + tap
    del eeg_d
    lap('features_s', t0)
    t0 = time.perf_counter()
    y = ops.target(audio_d, sfreq_audio)
    del audio_d
    y = y[20:-4]          # align the audio frames with the 20-frame context / 50 ms window of the features (train.py:144)
    medians, borders, q = ops.quantization(y, nb_intervals)
    lap('target_quantization_s', t0)

    t0 = time.perf_counter()
    if x_r is not None:
        if len(x_r) != len(y):
            # scipy.stats.spearmanr raises on unequal lengths; keep the reference's failure mode
            raise ValueError("all the input array dimensions must match: %d feature rows vs %d target rows" % (len(x_r), len(y)))
        rho_r, colsum_r = ops.spearman(x_r, y)
    else:
        rho_r, colsum_r = np.empty(0), np.empty(0)
    lap('spearman_s', t0)
    t0 = time.perf_counter()
    parts = _allgather_objects((np.asarray(rho_r), np.asarray(colsum_r)), group)
    cs = np.concatenate([p[0] for p in parts])
    colsum = np.concatenate([p[1] for p in parts])
    prof['rho_allgather_bytes'] = int(16 * len(cs)) if world > 1 else 0
    prof['rho_allgather_s'] = time.perf_counter() - t0
    cs[np.isclose(colsum, 0)] = 0
    select = np.argsort(np.abs(cs))[np.max([-nb_feats, -len(cs)]):]       # train.py:108: ascending |rho|, not index order

    # the selected columns, each from the rank that owns its channel
    t0 = time.perf_counter()
    n_rows = len(y)
    mine = np.nonzero((select >= 5 * c0) & (select < 5 * c1))[0]
    if world == 1:
        x_sel = ops.contiguous(x_r[:, ops.index(select, x_r)])
    else:
        x_sel = ops.zeros((n_rows, len(select)), like=y)
        if len(mine):
            x_sel[:, ops.index(mine, y)] = x_r[:, ops.index(select[mine] - 5 * c0, y)]
    del x_r
    lap('column_select_s', t0)
    prof['columns_allreduce_bytes'], prof['columns_allreduce_s'] = _allreduce_sum_(x_sel, group)

    minimum = min(len(x_sel), len(q))
    x_sel, q = x_sel[0:minimum, :], q[0:minimum, :]
    t0 = time.perf_counter()
    lo, hi = block_shard(minimum, rank, world)
    ident = np.arange(len(select))
    xl, ql = ops.contiguous(x_sel[lo:hi]), ops.contiguous(q[lo:hi])
    local_mean = ops.col_means(xl, ident) if hi > lo else np.zeros(len(select))
    mean_buf = np.concatenate([[float(hi - lo)], local_mean * (hi - lo)])
    b1, s1 = _allreduce_sum_(mean_buf, group)
    xbar = mean_buf[1:] / mean_buf[0]
    stats = ops.lda_stats(xl, ident, ql, nb_intervals, xbar)
    flat = np.concatenate([[stats['n']], stats['G'].ravel(), stats['sums'].ravel(), stats['counts'].ravel()])
    ops.sync()
    b2, s2 = _allreduce_sum_(flat, group)
    nf = len(select)
    o = 1
    stats['n'] = float(flat[0])
    stats['G'] = flat[o:o + nf * nf].reshape(nf, nf); o += nf * nf
    stats['sums'] = flat[o:o + stats['sums'].size].reshape(stats['sums'].shape); o += stats['sums'].size
    stats['counts'] = flat[o:o + stats['counts'].size].reshape(stats['counts'].shape)
    stats['xbar'] = xbar
    prof['stats_allreduce_bytes'], prof['stats_allreduce_s'] = b1 + b2, s1 + s2
    lap('lda_stats_s', t0)

    t0 = time.perf_counter()
    bins = list(range(rank, nb_mel_bins, world))
    fitted = fit_from_stats(stats, bins=bins)
    estimators = [None] * nb_mel_bins
    for part_bins, part in _allgather_objects((bins, fitted), group):
        for b, e in zip(part_bins, part):
            estimators[b] = e
    prof['eigen_solves_s'] = time.perf_counter() - t0
    prof['total_s'] = time.perf_counter() - t_all
    last_profile.clear()
    last_profile.update(prof)
    return x_sel, q, medians, estimators, select


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


def fit_from_stats(stats, tol=TOL, bins=None):
    """sklearn's _solve_svd (discriminant_analysis.py) restated on (G, class sums, counts): the SVD of the scaled,
    within-class-centred data matrix is replaced by the eigen-decomposition of its 150 x 150 Gram matrix.
    bins: the mel bins to fit (default all); the list returned follows their order."""
    N, xbar, G = stats['n'], stats['xbar'], stats['G']
    estimators = []
    for b in (range(stats['sums'].shape[0]) if bins is None else bins):
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
