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
