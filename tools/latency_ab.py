"""Real-time feed of the streaming chain under alternating scheduling policies inside ONE process (host noise on a shared box
drifts between runs: alternate short segments instead).  64-sample packets of 128 ch @ 2048 Hz, [segment_s] seconds per
segment, [rounds] rounds over the policies.  Usage: python tools/latency_ab.py [segment_s] [rounds]"""
import gc
import json
import os
import pickle
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class Policy:
    def __init__(self, name, fifo, pin):
        self.name, self.fifo, self.pin = name, fifo, pin

    def __enter__(self):
        self.aff = os.sched_getaffinity(0)
        self.pol, self.par = os.sched_getscheduler(0), os.sched_getparam(0)
        if self.pin:
            cores = sorted(self.aff)
            os.sched_setaffinity(0, {cores[len(cores) // 2]})
        if self.fifo:
            os.sched_setscheduler(0, os.SCHED_FIFO, os.sched_param(50))

    def __exit__(self, *a):
        os.sched_setscheduler(0, self.pol, self.par)
        os.sched_setaffinity(0, self.aff)


if __name__ == '__main__':
    seg = float(sys.argv[1]) if len(sys.argv) > 1 else 15.0
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    from livenodes import Node
    from sgs import synth
    import decode as dec_mod
    (W, b, cls), select, medians = bench.trained_model()
    ests = [bench._PlainEstimator(W[i], b[i], cls[i]) for i in range(40)]
    SR, N_CH, packet = bench.SR, bench.N_CH, 64
    policies = [Policy('time-sharing', False, False), Policy('SCHED_FIFO', True, False), Policy('SCHED_FIFO + pinned', True, True)]
    total = seg * rounds * len(policies) + 2.0
    x = synth.seeg_session(5, N_CH, SR, total)
    src = Node.Node(name='src', has_inputs=False)
    rec = dec_mod.setup_decoder(src, SR, pickle.dumps(ests), medians, [], select, gl_norm=10, packet_size=packet, include_soundcard=False)
    gl_node = rec[2].get_inputs()[0]
    lat, t_in = [], [0.0]
    gl_node.add_output(lambda f: lat.append(time.perf_counter() - t_in[0]))
    per_seg = int(seg * SR / packet)
    res = {p.name: [] for p in policies}
    gc.collect(); gc.disable()
    p_idx = 0
    for i in range(0, 64):                                   # warm-up packets, back to back
        src.output_data(np.array(x[p_idx * packet:(p_idx + 1) * packet])); p_idx += 1
    for r in range(rounds):
        for pol in policies:
            with pol:
                lat.clear()
                t_next = time.perf_counter()
                for _ in range(per_seg):
                    chunk = np.array(x[p_idx * packet:(p_idx + 1) * packet]); p_idx += 1
                    t_next += packet / SR
                    while time.perf_counter() < t_next:
                        time.sleep(0.0005)
                    t_in[0] = time.perf_counter()
                    src.output_data(chunk)
                L = np.array(lat) * 1e3
                res[pol.name].append({"frames": len(L), "p50": round(float(np.median(L)), 3), "p99": round(float(np.percentile(L, 99)), 3),
                                      "max": round(float(L.max()), 3), "over_1ms": int((L > 1.0).sum())})
    gc.enable()
    print(json.dumps(res))
