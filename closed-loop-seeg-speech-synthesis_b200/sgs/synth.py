"""Seeded synthetic sessions (sEEG + time-aligned audio) shared by the tests, the oracle runs and bench.py.

There are no recordings in the reference repository and no network here, so every workload is
synthetic: sEEG = broadband noise + 50/100/150 Hz line interference + a
high-gamma component whose power follows a slow "speech" envelope; audio = the same envelope
modulating a harmonic stack.  The envelope couples the two so that the 40 per-mel-bin LDA
problems are non-degenerate.
"""
import numpy as np


def _envelope(rng, duration_s, knot_hz=8.0):
    n_knots = int(np.ceil(duration_s * knot_hz)) + 2
    on = rng.random(n_knots) < 0.55
    level = np.where(on, 0.35 + 0.65 * rng.random(n_knots), 0.02 * rng.random(n_knots))
    return np.arange(n_knots) / knot_hz, level


def envelope_at(t, knots_t, knots_v):
    return np.interp(t, knots_t, knots_v)


def seeg_session(session, n_channels, sr, duration_s, dtype=np.float32):
    """(T x C) time-major sEEG, the layout the reference hands around (samples x channels)."""
    rng = np.random.default_rng(1000 + session)
    kt, kv = _envelope(np.random.default_rng(5000 + session), duration_s)
    n = int(round(duration_s * sr))
    t = np.arange(n) / float(sr)
    g = envelope_at(t, kt, kv)[:, None]
    phase = rng.uniform(0, 2 * np.pi, n_channels)[None, :]
    weight = rng.uniform(0.0, 1.0, n_channels)[None, :]
    x = 20.0 * rng.standard_normal((n, n_channels))
    x += 30.0 * np.sin(2 * np.pi * 50.0 * t[:, None] + phase)
    x += 10.0 * np.sin(2 * np.pi * 100.0 * t)[:, None]
    x += 5.0 * np.sin(2 * np.pi * 150.0 * t)[:, None]
    x += 60.0 * g * weight * rng.standard_normal((n, n_channels))
    return x.astype(dtype)


def seeg_sessions_device(sessions, n_channels, sr, duration_s, device='cuda', out=None):
    """(S x T x C) float32 torch tensor generated ON the device: the same signal model as seeg_session (same envelope knots,
    phases and channel weights per session number; the two Gaussian components come from torch's generator seeded with
    1000 + session, so the samples differ from the numpy ones while every statistic is the same).  Config-5 scale inputs
    (161 GB for 256 sessions) cannot be generated on the host and uploaded; they are made per rank where they are used."""
    import torch
    sessions = list(sessions)
    n = int(round(duration_s * sr))
    if out is None:
        out = torch.empty((len(sessions), n, n_channels), dtype=torch.float32, device=device)
    dev = out.device
    t = torch.arange(n, device=dev, dtype=torch.float64) / float(sr)
    lines = (10.0 * torch.sin(2 * np.pi * 100.0 * t) + 5.0 * torch.sin(2 * np.pi * 150.0 * t)).to(torch.float32)[:, None]
    tmp = torch.empty((n, n_channels), dtype=torch.float32, device=dev)
    for i, session in enumerate(sessions):
        rng = np.random.default_rng(1000 + session)
        kt, kv = _envelope(np.random.default_rng(5000 + session), duration_s)
        phase = torch.from_numpy(rng.uniform(0, 2 * np.pi, n_channels)).to(dev)[None, :]
        weight = torch.from_numpy(rng.uniform(0.0, 1.0, n_channels)).to(dev, torch.float32)[None, :]
        # envelope: linear interpolation between the 8 Hz knots (np.interp in seeg_session)
        knot_hz = 1.0 / (kt[1] - kt[0])
        pos = t * knot_hz
        i0 = pos.floor().long().clamp_(0, len(kv) - 2)
        kvd = torch.from_numpy(kv).to(dev)
        frac = (pos - i0.to(torch.float64))
        g = (kvd[i0] * (1.0 - frac) + kvd[i0 + 1] * frac).to(torch.float32)[:, None]
        gen = torch.Generator(device=dev)
        gen.manual_seed(1000 + session)
        x = out[i]
        x.normal_(0.0, 20.0, generator=gen)
        x += (30.0 * torch.sin(2 * np.pi * 50.0 * t[:, None] + phase)).to(torch.float32)
        x += lines
        tmp.normal_(0.0, 60.0, generator=gen)
        tmp *= g
        tmp *= weight
        x += tmp
    return out


def audio_session(session, duration_s, sr=16000):
    """Audio already at 16 kHz (the reference decimates 48 kHz by 3 before use, train.py:125).
    Includes the N(0, 1e-4) dither the reference adds at train.py:294."""
    rng = np.random.default_rng(2000 + session)
    kt, kv = _envelope(np.random.default_rng(5000 + session), duration_s)
    n = int(round(duration_s * sr))
    t = np.arange(n) / float(sr)
    g = envelope_at(t, kt, kv)
    f0 = 120.0 + 30.0 * np.sin(2 * np.pi * 0.7 * t)
    ph = 2 * np.pi * np.cumsum(f0) / sr
    a = np.zeros(n)
    for h, amp in enumerate([1.0, 0.6, 0.45, 0.3, 0.2, 0.15, 0.1, 0.08], start=1):
        a += amp * np.sin(h * ph + 0.3 * h)
    a = 0.15 * g * a + 0.004 * rng.standard_normal(n)
    a = np.clip(a, -0.5, 0.5)
    return a + rng.normal(0, 0.0001, n)


def audio_session_device(session, duration_s, sr=16000, device='cuda'):
    """audio_session generated on the device (float64 torch tensor): the same envelope knots, pitch contour and harmonic
    stack; the two noise terms come from torch's generator.  An hour at 48 kHz is 173 M samples - too slow to make on the
    host inside a benchmark."""
    import torch
    kt, kv = _envelope(np.random.default_rng(5000 + session), duration_s)
    n = int(round(duration_s * sr))
    t = torch.arange(n, device=device, dtype=torch.float64) / float(sr)
    knot_hz = 1.0 / (kt[1] - kt[0])
    pos = t * knot_hz
    i0 = pos.floor().long().clamp_(0, len(kv) - 2)
    kvd = torch.from_numpy(kv).to(device)
    frac = pos - i0.to(torch.float64)
    g = kvd[i0] * (1.0 - frac) + kvd[i0 + 1] * frac
    del pos, i0, frac
    f0 = 120.0 + 30.0 * torch.sin(2 * np.pi * 0.7 * t)
    ph = 2 * np.pi * torch.cumsum(f0, 0) / sr
    del f0, t
    a = torch.zeros(n, device=device, dtype=torch.float64)
    for h, amp in enumerate([1.0, 0.6, 0.45, 0.3, 0.2, 0.15, 0.1, 0.08], start=1):
        a += amp * torch.sin(h * ph + 0.3 * h)
    del ph
    gen = torch.Generator(device=device)
    gen.manual_seed(2000 + session)
    a = 0.15 * g * a + 0.004 * torch.randn(n, device=device, dtype=torch.float64, generator=gen)
    a.clamp_(-0.5, 0.5)
    return a + 0.0001 * torch.randn(n, device=device, dtype=torch.float64, generator=gen)


def logmel_utterances(n_utt, n_frames, medians, seed=3000):
    """Config-4 style input: log-mels drawn from each bin's quantisation medians (n_utt x T x bins)."""
    rng = np.random.default_rng(seed)
    nb, k = medians.shape
    idx = rng.integers(0, k, size=(n_utt, n_frames, nb))
    return medians[np.arange(nb)[None, None, :], idx]


def default_medians(n_bins=40, n_intervals=9, vmin=-16.0, vmax=-2.0):
    """Logistic representatives (local/quantization.py:105-107) for a typical log-mel range."""
    t = np.linspace(-9.5, 9.5, n_intervals, endpoint=True)
    lo = vmin + 0.05 * np.arange(n_bins)
    hi = vmax - 0.03 * np.arange(n_bins)
    L = np.abs(lo) + hi
    return L[:, None] / (1 + np.exp(-0.5 * t))[None, :] - np.abs(lo)[:, None]
