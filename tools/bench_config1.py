"""BASELINE config 1: offline decode of a synthetic 64-ch 1024 Hz 5-minute recording through decode.perform_offline_decoding
(host array in, host spectrogram + int16 audio out), wall clock, next to the CPU oracle on the same recording.
Usage: python tools/bench_config1.py [--no-cpu]"""
import json
import os
import pickle
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import decode  # noqa: E402
from sgs import synth  # noqa: E402
from sgs.synth import default_medians  # noqa: E402
from sgs.training import PackedLDA  # noqa: E402

if __name__ == '__main__':
    sr, n_ch, seconds = 1024, 64, 300.0
    rng = np.random.default_rng(11)
    W = rng.normal(0, 0.3, (40, 9, 150)); b = rng.normal(0, 1.0, (40, 9))
    cls = np.tile(np.arange(9, dtype=np.float64), (40, 1))
    select = rng.permutation(5 * n_ch)[:150].astype(np.int32)
    medians = default_medians(40, 9)
    ests = []
    for i in range(40):
        e = PackedLDA(); e.coef_, e.intercept_, e.classes_ = W[i], b[i], cls[i]
        ests.append(e)
    params = (pickle.dumps(ests), medians, [], select)
    x = synth.seeg_session(21, n_ch, sr, seconds).astype(np.float64)
    decode.perform_offline_decoding(params, x[: sr * 10], sr, 10)             # context, plans
    walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        spec, audio, _, _ = decode.perform_offline_decoding(params, x, sr, 10)
        walls.append(time.perf_counter() - t0)
    out = {"config": "64 ch x 1024 Hz x 300 s, decode.perform_offline_decoding (host in, host out)", "frames": int(len(spec)),
           "audio_samples": int(len(audio)), "wall_s": min(walls), "channel_seconds_per_s": n_ch * seconds / min(walls)}
    if '--no-cpu' not in sys.argv:
        import oracle as O
        noise = np.random.RandomState(1).rand(len(spec), 480)
        t0 = time.perf_counter()
        O.decode_streaming(x, sr, ests, select, medians, noise, gl_norm=10, chunk_size=32)
        out["cpu_oracle_wall_s"] = time.perf_counter() - t0
        out["speedup_vs_one_core"] = out["cpu_oracle_wall_s"] / out["wall_s"]
    print(json.dumps(out))
