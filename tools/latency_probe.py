"""Streaming latency of the node chain, fused (one sgs_chain_push per packet) against node-by-node device calls,
plus where a fused packet's time goes (host wall time of the C call; device time per kernel class).
Usage: python tools/latency_probe.py [seconds]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (puts the package on sys.path)

if __name__ == '__main__':
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    for fused in ('1', '0'):
        os.environ['SGS_FUSED_CHAIN'] = fused
        r = bench.latency_leg(seconds)
        r['fused'] = fused == '1'
        print(json.dumps(r), flush=True)
    os.environ['SGS_FUSED_CHAIN'] = '1'
    from sgs import chain, _lib
    calls = []
    orig = chain.FusedChain.push

    def timed(self, block, ends, idx):
        t0 = time.perf_counter()
        n = orig(self, block, ends, idx)
        calls.append((time.perf_counter() - t0, n))
        return n
    chain.FusedChain.push = timed
    r = bench.latency_leg(seconds)
    c = np.array([t for t, n in calls if n > 0][50:]) * 1e3
    print(json.dumps({"chain.push wall ms (incl. GL bookkeeping + noise draws)": {"p50": float(np.percentile(c, 50)),
                      "p99": float(np.percentile(c, 99))}, "packet p50 ms": r["packet_64"]["last_frame_of_packet"]["p50_ms"]}), flush=True)
    _lib.profile_enable(True)
    calls.clear()
    bench.latency_leg(5.0)
    out = {}
    for name in ('stream', 'lda', 'gl_blocks', 'gl_ola'):
        ms, n = _lib.profile_read(name)
        out[name] = {"launches": n, "mean_us": 1e3 * ms / max(n, 1)}
    print(json.dumps({"device time per launch": out}), flush=True)
