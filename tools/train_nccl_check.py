"""LDA training statistics sharded by rows over the ranks of one box (train.py:112-118 on N GPUs): every rank runs the
tensor-core statistics kernels on its shard, the mean and then (G, class sums, counts) are summed with NCCL all-reduces
(sgs/training.py:distributed_lda_stats), every rank fits the 40 estimators and the result is compared with one rank doing
everything.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_nccl_check.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
from sgs import _lib, training  # noqa: E402

if __name__ == '__main__':
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    _lib.ensure_init(local)
    n, width, nf, nb = 120000, 640, 150, 40
    rng = np.random.default_rng(5)                               # every rank draws the same data and takes its slice
    X = rng.normal(9.0, 0.6, (n, width))
    labels = rng.integers(0, 9, (n, nb)).astype(np.float64)
    for b in range(nb):
        X[np.arange(n), (labels[:, b].astype(int) * 7 + b) % width] += 0.7
    select = np.sort(rng.permutation(width)[:nf])
    cuts = np.linspace(0, n, world + 1).astype(int)
    lo, hi = cuts[rank], cuts[rank + 1]
    t0 = time.perf_counter()
    stats = training.distributed_lda_stats(X[lo:hi], select, labels[lo:hi])
    ests = training.fit_from_stats(stats)
    dt = time.perf_counter() - t0
    if rank == 0:
        full = training.lda_stats(X, select, labels)
        ref = training.fit_from_stats(full)
        relG = float(np.abs(stats['G'] - full['G']).max() / np.abs(full['G']).max())
        relC = max(float(np.abs(a.coef_ - b.coef_).max() / np.abs(b.coef_).max()) for a, b in zip(ests, ref))
        Xt = rng.normal(9.0, 0.7, (20000, width))[:, select]
        flips = int(sum((a.predict(Xt) != b.predict(Xt)).sum() for a, b in zip(ests, ref)))
        ok = stats['n'] == n and np.array_equal(stats['counts'], full['counts']) and relG < 1e-12 and relC < 1e-8 and flips == 0
        print(json.dumps({"world": world, "backend": dist.get_backend() if world > 1 else None, "rows": n, "n_total": stats['n'],
                          "counts_equal": bool(np.array_equal(stats['counts'], full['counts'])), "max_rel_diff_G": relG,
                          "max_rel_diff_coef": relC, "prediction_flips_of_800000": flips, "seconds_sharded_fit": dt,
                          "verdict": "OK" if ok else "MISMATCH"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
