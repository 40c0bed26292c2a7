"""BASELINE config 3: train.train on a synthetic 128-channel session (default 1 h @ 2048 Hz sEEG, 48 kHz audio), host arrays in,
fitted model out, with the per-stage wall times train.py logs.  Usage: python tools/bench_train_e2e.py [seconds] [channels]"""
import json
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
import train  # noqa: E402
from sgs import synth, _lib  # noqa: E402

if __name__ == '__main__':
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    n_ch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    sr = 2048
    stages = {}

    class Grab(logging.Handler):
        def emit(self, record):
            m = record.getMessage()
            if m.startswith('Finished method ['):
                name = m.split('[')[1].split(']')[0]
                stages[name] = float(m.split(' in ')[1].split(' ')[0])
    logging.getLogger('utils.py').addHandler(Grab())
    logging.getLogger('utils.py').setLevel(logging.INFO)
    _lib.ensure_init(0)
    t0 = time.perf_counter()
    rng = np.random.default_rng(1)
    audio = synth.audio_session(1, seconds, 48000)
    env = np.abs(audio[::48000 // sr][:int(seconds * sr)])
    env = np.convolve(env, np.ones(256) / 256, mode='same')
    eeg = rng.standard_normal((int(seconds * sr), n_ch), dtype=np.float32).astype(np.float64) * 50.0
    eeg[:, ::3] *= (1.0 + 4.0 * env / env.max())[:, None]                 # a third of the channels carry the speech envelope
    t_gen = time.perf_counter() - t0
    train.train(eeg[: sr * 20], audio[: 48000 * 20], sr, 48000, [])        # warm-up: context, plans, allocator
    walls = []
    for _ in range(2):                                                     # the first full-size call also grows the memory pools
        stages.clear()
        t0 = time.perf_counter()
        x_train, q, medians, estimators, select = train.train(eeg, audio, sr, 48000, [])
        walls.append(time.perf_counter() - t0)
    dt = walls[-1]
    print(json.dumps({"seconds_of_data": seconds, "channels": n_ch, "rows": int(x_train.shape[0]), "train_wall_s": dt, "first_call_wall_s": walls[0],
                      "channel_seconds_per_s": n_ch * seconds / dt, "stages_s": stages, "data_generation_s": t_gen,
                      "classes_per_bin": [int(len(e.classes_)) for e in estimators][:8]}))
