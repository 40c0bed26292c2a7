"""Host-side training logic on CPU: the LDA fit from sufficient statistics against scikit-learn, and the
row-sharded all-reduce path with two gloo ranks."""
import os
import sys

import numpy as np
import pytest

from sgs import training


def numpy_stats(X, labels, n_classes=9, xbar=None):
    xbar = X.mean(0) if xbar is None else xbar
    Xc = X - xbar
    nb = labels.shape[1]
    sums = np.zeros((nb, n_classes, X.shape[1])); counts = np.zeros((nb, n_classes))
    for b in range(nb):
        for k in range(n_classes):
            m = labels[:, b] == k
            sums[b, k] = Xc[m].sum(0); counts[b, k] = m.sum()
    return dict(n=float(len(X)), xbar=xbar, G=Xc.T @ Xc, sums=sums, counts=counts)


def make_problem(seed=0, n=4000, f=30, bins=5):
    rng = np.random.default_rng(seed)
    X = rng.normal(9.0, 0.5, (n, f))
    labels = np.stack([rng.integers(0, 9, n), rng.choice([0, 2, 3, 7], n), rng.choice([1, 5], n),
                       rng.integers(0, 9, n), rng.integers(0, 6, n)], axis=1).astype(float)[:, :bins]
    for b in range(bins):
        X[np.arange(n), (labels[:, b].astype(int) * 3 + b) % f] += 0.8
    return X, labels


def test_fit_from_stats_matches_sklearn():
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    X, labels = make_problem()
    ests = training.fit_from_stats(numpy_stats(X, labels))
    Xt = np.random.default_rng(9).normal(9.0, 0.6, (3000, X.shape[1]))
    for b, e in enumerate(ests):
        ref = LinearDiscriminantAnalysis().fit(X, labels[:, b])
        assert np.array_equal(e.classes_, ref.classes_)
        assert np.abs(e.coef_ - ref.coef_).max() <= 1e-8 * np.abs(ref.coef_).max()
        assert np.abs(e.intercept_ - ref.intercept_).max() <= 1e-8 * np.abs(ref.intercept_).max()
        assert np.array_equal(e.predict(Xt), ref.predict(Xt))


WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import torch.distributed as dist
from sgs import training
from test_train_host import make_problem, numpy_stats
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
rank, world = dist.get_rank(), dist.get_world_size()
X, labels = make_problem()
cut = [0, 1700, len(X)]
Xl, Ll = X[cut[rank]:cut[rank + 1]], labels[cut[rank]:cut[rank + 1]]
xbar = training.global_mean(Xl.mean(0), len(Xl))
stats = training.allreduce_stats(numpy_stats(Xl, Ll, xbar=xbar))
ests = training.fit_from_stats(stats)
np.savez(sys.argv[3] + '.%d.npz' % rank, n=stats['n'], xbar=xbar, **{'coef%d' % b: e.coef_ for b, e in enumerate(ests)})
dist.destroy_process_group()
"""


def test_two_rank_gloo_allreduce_fit(tmp_path):
    """Row shards on two ranks -> all-reduced mean and statistics -> every rank fits the same model as one process."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'closed-loop-seeg-speech-synthesis_b200')
    out = str(tmp_path / 'w')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29533', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script), pkg, os.path.join(root, 'tests'), out], env=dict(env, RANK=str(r)))
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    X, labels = make_problem()
    single = training.fit_from_stats(numpy_stats(X, labels))
    for r in range(2):
        R = np.load(out + '.%d.npz' % r)
        assert float(R['n']) == len(X)
        assert np.abs(R['xbar'] - X.mean(0)).max() < 1e-12
        for b, e in enumerate(single):
            assert np.abs(R['coef%d' % b] - e.coef_).max() <= 1e-9 * np.abs(e.coef_).max()


SHARD_WORKER = r"""
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import decode
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
out = {}
for n in (0, 1, 2, 7, 256):
    lo, hi = decode.session_shard(n)                       # rank / world from the process group
    t = torch.tensor([lo, hi])
    both = [torch.zeros(2, dtype=torch.long) for _ in range(dist.get_world_size())]
    dist.all_gather(both, t)
    out[str(n)] = [b.tolist() for b in both]
# the bench's timing rule: the step time of the job is the MAX over ranks
ms = torch.tensor([10.0 + dist.get_rank()], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
out['max_ms'] = float(ms.item())
if dist.get_rank() == 0:
    json.dump(out, open(sys.argv[2], 'w'))
dist.destroy_process_group()
"""


def test_two_rank_gloo_session_sharding(tmp_path):
    """Decode shards by session with no data-path collective: the ranks' slices partition the job exactly."""
    import json
    import subprocess
    import decode
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'closed-loop-seeg-speech-synthesis_b200')
    script = tmp_path / 'shard_worker.py'
    script.write_text(SHARD_WORKER)
    out = str(tmp_path / 'shards.json')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29534', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script), pkg, out], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    R = json.load(open(out))
    for n in (0, 1, 2, 7, 256):
        (lo0, hi0), (lo1, hi1) = R[str(n)]
        assert lo0 == 0 and hi0 == lo1 and hi1 == n and abs((hi0 - lo0) - (hi1 - lo1)) <= 1
    assert R['max_ms'] == 11.0
    assert [decode.session_shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    with pytest.raises(ValueError):
        decode.session_shard(4, 2, 2)


class NumpyOps:
    """CPU restatements of the device operators of sgs/training.py:DeviceOps (oracle / scipy / numpy): they let the SHARDED
    orchestration of train.train - channel-block features + Spearman, all-gathered correlations, summed column matrix,
    row-sharded statistics, dealt eigen-solves - run on CPU ranks over gloo."""

    def upload(self, a, dtype=None):
        return np.ascontiguousarray(a) if dtype is None else np.ascontiguousarray(a, dtype=dtype)

    def sync(self):
        pass

    def features(self, eeg, sfreq_eeg):
        import oracle as O
        return O.herff2016_b(np.asarray(eeg, dtype=np.float64), sfreq_eeg)

    def target(self, audio, audio_sr):
        import oracle as O
        assert audio_sr == 16000
        return O.compute_spectrogram(audio, 16000, 0.016, 0.01)

    def quantization(self, y, nb_intervals):
        import oracle as O
        medians, borders = O.compute_borders_logistic(y, nb_intervals)
        return medians, borders, O.quantize_spectrogram(y, borders)

    def spearman(self, x, y):
        from scipy.stats import spearmanr
        tgt = np.mean(y, axis=1)
        return np.array([spearmanr(x[:, f], tgt)[0] for f in range(x.shape[1])]), x.sum(axis=0)

    def zeros(self, shape, like):
        return np.zeros(shape)

    def index(self, idx, like):
        return np.asarray(idx, dtype=np.int64)

    def contiguous(self, a):
        return np.ascontiguousarray(a)

    def to_host(self, a):
        return a

    def col_means(self, x, select):
        return x[:, select].mean(axis=0)

    def lda_stats(self, x, select, labels, n_classes, xbar):
        return numpy_stats(x[:, select], labels, n_classes, xbar)


def sharded_problem():
    from sgs import synth
    sr, n_ch, dur = 1024, 10, 12.0
    return synth.seeg_session(31, n_ch, sr, dur).astype(np.float64), synth.audio_session(31, dur), sr


SHARDED_WORKER = r"""
import os, sys, pickle, numpy as np
for p in sys.argv[1:4]:
    sys.path.insert(0, p)
import torch.distributed as dist
from sgs import training
from test_train_host import NumpyOps, sharded_problem
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
eeg, audio, sr = sharded_problem()
x, q, medians, est, select = training.sharded_fit(eeg, audio, sr, 16000, nb_mel_bins=40, nb_feats=20, ops=NumpyOps())
with open(sys.argv[4] + '.%d.pkl' % dist.get_rank(), 'wb') as fh:
    pickle.dump(dict(x=x, q=q, medians=medians, select=select, coef=[e.coef_ for e in est], icpt=[e.intercept_ for e in est],
                     cls=[e.classes_ for e in est], prof=dict(training.last_profile)), fh)
dist.destroy_process_group()
"""


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_train_over_gloo_equals_single_rank(tmp_path, world):
    """train.train's multi-GPU plan on CPU ranks: every rank ends with the single-rank model - `select` identical, the
    selected feature matrix identical to the bit (each column comes from exactly one rank), coefficients to round-off of
    the re-ordered sums, and not one prediction differs.  3 ranks: 10 channels and 40 bins do not divide evenly."""
    import pickle
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'closed-loop-seeg-speech-synthesis_b200')
    out = str(tmp_path / 'w')
    script = tmp_path / 'sharded_worker.py'
    script.write_text(SHARDED_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=str(29540 + world), WORLD_SIZE=str(world), OMP_NUM_THREADS='2')
    procs = [subprocess.Popen([sys.executable, str(script), pkg, os.path.join(root, 'tests'), os.path.join(root, 'oracle'), out],
                              env=dict(env, RANK=str(r))) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    eeg, audio, sr = sharded_problem()
    x1, q1, med1, est1, sel1 = training.sharded_fit(eeg, audio, sr, 16000, nb_mel_bins=40, nb_feats=20, ops=NumpyOps())
    assert training.last_profile['world'] == 1 and training.last_profile['stats_allreduce_bytes'] == 0
    # the single-rank orchestration is the reference's train (oracle.train restates train.py:132-168)
    import oracle as O
    xo, qo, medo, esto, selo = O.train(eeg, audio, sr, [], nb_feats=20)
    assert np.array_equal(sel1, selo) and np.array_equal(x1, xo) and np.array_equal(q1, qo) and np.array_equal(med1, medo)
    for r in range(world):
        R = pickle.load(open(out + '.%d.pkl' % r, 'rb'))
        assert R['prof']['world'] == world and R['prof']['stats_allreduce_bytes'] > 8 * (20 * 20 + 40 * 9 * 20 + 40 * 9)
        assert R['prof']['columns_allreduce_bytes'] == x1.size * 8
        assert np.array_equal(R['select'], sel1)
        assert np.array_equal(R['x'], x1) and np.array_equal(R['q'], q1) and np.array_equal(R['medians'], med1)
        for b in range(40):
            assert np.array_equal(R['cls'][b], est1[b].classes_)
            assert np.abs(R['coef'][b] - est1[b].coef_).max() <= 1e-8 * np.abs(est1[b].coef_).max()
            scores = x1 @ R['coef'][b].T + R['icpt'][b]
            pred = R['cls'][b][(scores[:, 0] > 0).astype(int)] if scores.shape[1] == 1 else R['cls'][b][scores.argmax(1)]
            assert np.array_equal(pred, est1[b].predict(x1))
            assert np.array_equal(est1[b].predict(x1), esto[b].predict(xo))
