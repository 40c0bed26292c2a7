"""Modal tail of the feature cascade (sgs/modal.py) on CPU: the closed-form expansion of the impulse response of all filter
states, and the start state the device's tail kernel forms from it, against scipy.signal.sosfilt run over the whole warm-up
(what the reference's filter state holds after a long recording, local/offline.py:63-97)."""
import numpy as np
import pytest
import scipy.signal

from sgs import design
from sgs.modal import BLOCK, MODE_GROUP, TAIL_WARPS, ModalTail, cascade_step_matrix, modal_expansion


@pytest.mark.parametrize('sr,ln', [(2048, 50), (1024, 60)])
def test_modal_expansion_reproduces_the_impulse_response_of_every_state(sr, ln):
    p = design.FeaturePlan(sr, line_noise=ln)
    lam, gamma = modal_expansion(p.coef)
    assert np.all(lam.imag > 0) and np.all(np.abs(lam) < 1)
    A, B = cascade_step_matrix(p.coef)
    v = B.copy()
    for n in range(4001):
        if n in (50, 500, 4000):
            # (the first samples cancel between modes of amplitude 1e6: only n >> 1 is ever used)
            approx = 2.0 * np.real(gamma @ lam ** n)
            assert np.max(np.abs(approx - v)) < 1e-12 * np.abs(v).max()
        v = A @ v


@pytest.mark.parametrize('sr,ln', [(2048, 50), (1024, 50), (2048, 60)])
def test_tail_then_short_cascade_gives_the_state_of_the_long_run(sr, ln):
    p = design.FeaturePlan(sr, line_noise=ln)
    mt = ModalTail(p.coef, 2.0 ** -50)
    assert mt.n_modes % MODE_GROUP == 0 and 0 < mt.n_modes <= 16 and mt.near_len % 64 == 0 and mt.far_len % BLOCK == 0
    assert mt.cost < 0.3 * 99 * mt.horizon                                # the point of it: a fraction of the cascade's work
    # every block of every group is dealt to exactly one warp
    for g in range(mt.n_modes // MODE_GROUP):
        seen = np.zeros(int(mt.mode_len[g * MODE_GROUP]) // BLOCK, dtype=int)
        for w in range(TAIL_WARPS):
            seen[mt.warp_blocks[w, g, 0]:mt.warp_blocks[w, g, 1]] += 1
        assert np.all(seen == 1)
    rng = np.random.default_rng(sr + ln)
    T = mt.horizon + 3000
    t = np.arange(T)
    x = 40.0 * rng.standard_normal(T) + 25.0 * np.sin(2 * np.pi * ln * t / sr) + 8.0 * np.sin(2 * np.pi * 2 * ln * t / sr + 1.0)
    sos = np.vstack(p.filters)
    zero = np.zeros((sos.shape[0], 2))
    _, full = scipy.signal.sosfilt(sos, x, zi=zero)                       # the state after the whole run
    scale = 40.0 * mt.state_scale.reshape(-1, 2)
    t_near = T - mt.near_len
    for start in (mt.reference_state(x[:t_near]), mt.kernel_state(x[:t_near])):
        _, got = scipy.signal.sosfilt(sos, x[t_near:], zi=start.reshape(-1, 2))
        assert np.max(np.abs(got - full) / scale) < 2e-12
    # and the tail matters: the short cascade alone is off by orders of magnitude more
    _, short = scipy.signal.sosfilt(sos, x[t_near:], zi=zero)
    assert np.max(np.abs(short - full) / scale) > 1e-6


def test_tail_near_the_start_of_a_recording_takes_the_rest_from_the_initial_state():
    """A piece that starts inside the tail's horizon: the samples there are, plus lambda^t (kappa . s_init) for the
    reference's cold-start state (each filter's sosfilt_zi times its first input, local/offline.py:39-62)."""
    sr, ln = 2048, 50
    p = design.FeaturePlan(sr, line_noise=ln)
    mt = ModalTail(p.coef, 2.0 ** -50)
    sos = np.vstack(p.filters)
    rng = np.random.default_rng(11)
    for T in (2 * mt.near_len, mt.horizon // 2 // 64 * 64, mt.horizon - 64):
        x = 40.0 * rng.standard_normal(T) + 500.0 + 30.0 * np.sin(2 * np.pi * ln * np.arange(T) / sr)
        zi, v = [], x[0]
        for f in p.filters:
            z = scipy.signal.sosfilt_zi(f) * v
            zi.append(z)
            v = scipy.signal.sosfilt(f, [v], zi=z)[0][0]
        zi = np.vstack(zi)
        _, full = scipy.signal.sosfilt(sos, x, zi=zi)
        t_near = T - mt.near_len
        scale = 40.0 * mt.state_scale.reshape(-1, 2)
        for start in (mt.reference_state(x[:t_near], zi.ravel()), mt.kernel_state(x[:t_near], zi.ravel())):
            _, got = scipy.signal.sosfilt(sos, x[t_near:], zi=start.reshape(-1, 2))
            assert np.max(np.abs(got - full) / scale) < 2e-12
        # without the initial state's share the DC offset of the recording is missing from the slow modes
        _, got = scipy.signal.sosfilt(sos, x[t_near:], zi=mt.reference_state(x[:t_near]).reshape(-1, 2))
        assert T > mt.horizon - 128 or np.max(np.abs(got - full) / scale) > 1e-9


def test_tail_under_adverse_inputs():
    """A line-noise harmonic 100 times the broadband level sits exactly on the notch zeros: the states of the later sections
    barely see it while the single modal sums are large - the worst case for the tail, and still three orders below the
    1e-9 the features are held to.  Errors relative to each state's rms under that very input."""
    p = design.FeaturePlan(2048)
    mt = ModalTail(p.coef, 2.0 ** -50)
    sos = np.vstack(p.filters)
    zero = np.zeros((sos.shape[0], 2))
    rng = np.random.default_rng(5)
    T = mt.horizon + 8192
    t = np.arange(T)
    for x in (5.0 * rng.standard_normal(T) + 500.0 * np.sin(2 * np.pi * 100 * t / 2048),
              1e4 + 50.0 * rng.standard_normal(T) + 3000.0 * (t > T // 2)):
        _, full = scipy.signal.sosfilt(sos, x, zi=zero)
        t_near = T - mt.near_len
        _, got = scipy.signal.sosfilt(sos, x[t_near:], zi=mt.kernel_state(x[:t_near]).reshape(-1, 2))
        states, zi = [], scipy.signal.sosfilt(sos, x[:T - 4096], zi=zero)[1]
        for k in range(0, 4096, 64):
            zi = scipy.signal.sosfilt(sos, x[T - 4096 + k:T - 4096 + k + 64], zi=zi)[1]
            states.append(zi.copy())
        rms = np.sqrt(np.mean(np.array(states) ** 2, axis=0))
        assert np.max(np.abs(got - full) / rms) < 1e-11
