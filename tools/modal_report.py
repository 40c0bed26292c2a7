"""The modal tail of the feature warm-up (sgs/modal.py) for every filter configuration: split point, modes kept, their
horizons, cost against the zero-state warm-up, and the error of tail + short cascade run against scipy.signal.sosfilt over the
whole warm-up on a seeded input (CPU only).  Usage: python tools/modal_report.py"""
import json
import os
import sys

import numpy as np
import scipy.signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
from sgs.design import FeaturePlan  # noqa: E402
from sgs.modal import CASCADE_OPS, ModalTail  # noqa: E402

if __name__ == '__main__':
    for sr, ln in ((2048, 50), (1024, 50), (2048, 60), (1024, 60)):
        p = FeaturePlan(sr, line_noise=ln)
        mt = ModalTail(p.coef, 2.0 ** -50)
        rng = np.random.default_rng(1)
        T = mt.horizon + 4096
        x = 50.0 * rng.standard_normal(T) + 30.0 * np.sin(2 * np.pi * ln * np.arange(T) / sr)
        sos = np.vstack(p.filters)
        zero = np.zeros((sos.shape[0], 2))
        _, full = scipy.signal.sosfilt(sos, x, zi=zero)
        t_near = T - mt.near_len
        _, got = scipy.signal.sosfilt(sos, x[t_near:], zi=mt.kernel_state(x[:t_near]).reshape(-1, 2))
        err = float(np.max(np.abs(got - full) / (50.0 * mt.state_scale.reshape(-1, 2))))
        print(json.dumps({"sample_rate_hz": sr, "line_noise_hz": ln, "near_len": mt.near_len, "modes": mt.n_modes,
                          "pole_radii": [round(float(abs(v)), 5) for v in mt.lam if v != 0],
                          "mode_len": sorted(set(int(v) for v in mt.mode_len), reverse=True), "horizon": mt.horizon,
                          "ops_per_cut": mt.cost, "ops_per_cut_zero_state_same_horizon": CASCADE_OPS * mt.horizon,
                          "state_error_rel_to_scale": err}))
