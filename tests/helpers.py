"""Shared helpers for the parity tests (inputs regenerated from seeds, noise replay)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def node_noise(seed, n_frames, block=480, first=1):
    """Replays the np.random.rand(block) draws of GriffinLimSynthesis (GriffinLim.py:90): one per frame from `first` on."""
    rs = np.random.RandomState(seed)
    noise = np.zeros((n_frames, block))
    for k in range(first, n_frames):
        noise[k] = rs.rand(block)
    return noise


def batch_noise(seed, n_frames, win=800, hop=160, n_bins=401):
    """np.random.rand(2*T*n_bins) of offline.griffin_lim (offline.py:164); only the head matters (R4)."""
    rs = np.random.RandomState(seed)
    return rs.rand(2 * n_frames * n_bins)[:hop * (n_frames - 1) + win]


def model128():
    """The reference-trained model of the 128-channel configurations (tests/golden/model128.npz, written by
    oracle/gen_golden.py:gen_model128 from the unmodified reference's train.train on 120 s of sgs.synth session 101):
    ((W, bias, classes), select, medians, fixture)."""
    G = load('model128.npz')
    return (G['coef'], G['intercept'], G['classes']), G['select'], G['medians'], G
