"""Which packets of the streaming latency leg are slow: wall time of every fused chain push (64-sample packets), with the
indices and spacing of the slow ones.  Usage: python tools/latency_outliers.py [seconds] [repeats]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == '__main__':
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    from sgs import chain
    calls = []
    orig = chain.FusedChain.push

    def timed(self, block, ends, idx):
        t0 = time.perf_counter()
        n = orig(self, block, ends, idx)
        calls.append((t0, time.perf_counter() - t0, n))
        return n
    chain.FusedChain.push = timed
    for rep in range(repeats):
        calls.clear()
        r = bench.latency_leg(seconds)
        c = [(t0, dt * 1e3, n) for t0, dt, n in calls]
        half = len(c) // 2                                   # first half: 64-sample packets, second half: 32-sample packets
        for name, part in (('packet_64', [v for v in c if True][:960]), ('packet_32', c[960:])):
            dt = np.array([v[1] for v in part][50:])
            t0 = np.array([v[0] for v in part][50:])
            slow = np.nonzero(dt > 1.6 * np.median(dt))[0]
            print(json.dumps({"rep": rep, "leg": name, "pushes": len(dt), "p50_ms": float(np.median(dt)), "p99_ms": float(np.percentile(dt, 99)),
                              "slow": len(slow), "slow_idx": slow[:40].tolist(), "slow_ms": np.round(dt[slow[:40]], 3).tolist(),
                              "slow_gap_ms": np.round(np.diff(t0[slow[:40]]) * 1e3, 1).tolist(),
                              "frame_p99_ms": r[name]["p99_ms"]}), flush=True)
