"""When do the slow frames of the real-time feed happen, and where is the time spent?  Feeds 64-sample packets of 128 ch @ 2048 Hz
in real time for [seconds] through the decode.setup_decoder graph and lists every frame over 1 ms: wall-clock time since start,
latency, time inside the C chain push, device time of the push's kernels (profile classes).  Usage: python tools/latency_events.py [seconds]"""
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

if __name__ == '__main__':
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
    from livenodes import Node
    from sgs import synth, chain, _lib
    import decode as dec_mod
    (W, b, cls), select, medians = bench.trained_model()
    ests = [bench._PlainEstimator(W[i], b[i], cls[i]) for i in range(40)]
    SR, N_CH, packet = bench.SR, bench.N_CH, 64
    x = synth.seeg_session(5, N_CH, SR, seconds + 1.0)
    src = Node.Node(name='src', has_inputs=False)
    rec = dec_mod.setup_decoder(src, SR, pickle.dumps(ests), medians, [], select, gl_norm=10, packet_size=packet, include_soundcard=False)
    gl_node = rec[2].get_inputs()[0]
    lat, t_in, push = [], [0.0], []
    gl_node.add_output(lambda f: lat.append((time.perf_counter(), time.perf_counter() - t_in[0])))
    orig = chain.FusedChain.push
    _lib.ensure_init(0)
    _lib.profile_enable(True)
    classes = ('stream', 'lda', 'gl_blocks', 'gl_ola')

    def timed(self, block, ends, idx):
        before = [_lib.profile_read(c)[0] for c in classes]
        t0 = time.perf_counter()
        r = orig(self, block, ends, idx)
        dt = time.perf_counter() - t0
        push.append((t0, dt, sum(_lib.profile_read(c)[0] - v for c, v in zip(classes, before))))
        return r
    chain.FusedChain.push = timed
    n_packets = int(seconds * SR / packet)
    gc.collect(); gc.disable()
    t_start = t_next = time.perf_counter()
    for p in range(n_packets):
        chunk = np.array(x[p * packet:(p + 1) * packet])
        t_next += packet / SR
        while time.perf_counter() < t_next:
            time.sleep(0.0005)
        t_in[0] = time.perf_counter()
        src.output_data(chunk)
    gc.enable()
    L = np.array([v[1] for v in lat[30:]]) * 1e3
    slow = [(round(t - t_start, 3), round(d * 1e3, 3)) for t, d in lat[30:] if d > 1e-3]
    P = np.array([v[1] for v in push[10:]]) * 1e3
    slow_push = [(round(t0 - t_start, 3), round(dt * 1e3, 3), round(dev, 3)) for t0, dt, dev in push[10:] if dt > 1e-3]
    print(json.dumps({"frames": len(L), "p50_ms": float(np.median(L)), "p99_ms": float(np.percentile(L, 99)), "max_ms": float(L.max()),
                      "frames_over_1ms_at_s_and_ms": slow, "push_p50_ms": float(np.median(P)),
                      "pushes_over_1ms_at_s_ms_devicems": slow_push}))
