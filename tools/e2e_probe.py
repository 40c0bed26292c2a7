"""Where the end-to-end (host in -> host out) decode time goes: pinned H2D bandwidth, resident decode time per batch
size, and OfflineDecoder.decode on pinned host sessions for several sessions_per_batch.
Usage: python tools/e2e_probe.py [sessions]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import decode as dec_mod  # noqa: E402
from sgs import _lib  # noqa: E402

if __name__ == '__main__':
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    _lib.ensure_init(0)
    torch.cuda.set_device(0)
    rng = np.random.default_rng(7)
    model, select, medians = bench.trained_model()
    decoder = dec_mod.OfflineDecoder(model, medians, select, bench.SR, gl_norm=10, packet_size=64)
    T = int(bench.SR * bench.DUR)
    xh = torch.empty((S, T, bench.N_CH), dtype=torch.float32).pin_memory()
    xh.normal_(0, 50.0)
    out = {}
    d = torch.empty((T, bench.N_CH), dtype=torch.float32, device='cuda')
    for _ in range(2):
        d.copy_(xh[0], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(S):
        d.copy_(xh[s], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["h2d_pinned_GBps"] = S * xh[0].numel() * 4 / dt / 1e9
    out["h2d_ms_per_session"] = dt / S * 1e3
    for nb in (1, 2, 4):
        if nb > S:
            break
        xd = xh[:nb].cuda()
        for _ in range(2):
            r = decoder.decode(xd, None, 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            r = decoder.decode(xd, None, 1)
        torch.cuda.synchronize()
        out["resident_ms_per_session_batch%d" % nb] = (time.perf_counter() - t0) / 3 / nb * 1e3
        del xd, r
    xn = xh.numpy()
    for pb in (1, 2, 4):
        if pb >= S:
            break
        for _ in range(2):
            decoder.decode(xn, None, 1, sessions_per_batch=pb, pinned_outputs=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            decoder.decode(xn, None, 1, sessions_per_batch=pb, pinned_outputs=True)
        torch.cuda.synchronize()
        out["e2e_ms_per_session_per_batch%d" % pb] = (time.perf_counter() - t0) / 3 / S * 1e3
    print(json.dumps(out))
