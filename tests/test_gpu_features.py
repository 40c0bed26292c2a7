"""CUDA feature extraction (through the C ABI) against the oracle and the reference-generated fixtures."""
import numpy as np
import pytest

import oracle as O
from sgs import synth
from sgs.features import FeatureExtractor
from helpers import load

pytestmark = pytest.mark.gpu

TOL = 1e-9      # absolute, on log-power values of magnitude ~5-12 (fp64 recurrences; FMA contraction differs from scipy)


@pytest.mark.parametrize('sr', [1024, 2048])
@pytest.mark.parametrize('ln', [50, 60])
def test_offline_against_reference_fixture(sr, ln):
    G = load('features.npz')
    x = synth.seeg_session(7, 6, sr, 1.37)
    fe = FeatureExtractor(sr, line_noise=ln)
    lp = fe.log_power(x.astype(np.float64))
    key = 'sr%d_ln%d' % (sr, ln)
    assert lp.shape == G[key + '_offline_nostack'].shape
    assert np.abs(lp - G[key + '_offline_nostack']).max() < TOL
    st = fe.stack(lp)
    assert st.shape == G[key + '_offline'].shape
    assert np.abs(st - G[key + '_offline']).max() < TOL
    # fp32 input path gives the same result (the synthetic data is exactly representable in fp32)
    assert np.array_equal(fe.log_power(x), lp)


@pytest.mark.parametrize('sr', [1024, 2048])
@pytest.mark.parametrize('ln', [50, 60])
def test_online_against_reference_fixture(sr, ln):
    G = load('features.npz')
    x = synth.seeg_session(7, 6, sr, 1.37)
    fe = FeatureExtractor(sr, line_noise=ln, frame_len_ms=50, frame_shift_ms=10)
    for cs, p in ((32, 16), (64, 64)):
        g = G['sr%d_ln%d_online_cs%d_p%d' % (sr, ln, cs, p)]
        st = fe.stack(fe.log_power(x, online=True, chunk_size=cs), online=True)
        assert st.shape == g.shape
        assert np.abs(st - g).max() < TOL


def test_float64_recording_keeps_its_bits():
    """A float64 recording that is NOT representable in float32 (the reference hands float64 arrays around): the offline
    kernels and the streaming kernel filter it at full precision.  (Until round 2 the offline loader narrowed every sample
    to float32 first, which no test saw because the synthetic sessions are float32 numbers.)"""
    sr = 1024
    rng = np.random.default_rng(3)
    x = synth.seeg_session(9, 5, sr, 1.2).astype(np.float64)
    x = x * (1.0 + 1e-5 * rng.standard_normal(x.shape))
    assert np.abs(x - x.astype(np.float32)).max() > 1e-8 * np.abs(x).max()
    fe = FeatureExtractor(sr)
    want = O.herff2016_b(x, sr, skip_stacking=True)
    assert np.abs(fe.log_power(x) - want).max() < TOL
    narrowed = O.herff2016_b(x.astype(np.float32).astype(np.float64), sr, skip_stacking=True)
    assert np.abs(narrowed - want).max() > 30 * TOL                       # what the narrowing cost (6.6e-8 in log-power)
    on = O.ecog_feat_calc(x, sr, 50, 10, 4, 5, 50, 32, stacked=False)
    assert np.abs(fe.log_power(x, online=True, chunk_size=32) - on).max() < TOL


def test_short_and_empty():
    G = load('features.npz')
    x = synth.seeg_session(8, 1, 1024, 0.30)
    fe = FeatureExtractor(1024)
    lp = fe.log_power(x)
    assert np.abs(lp - G['short_offline_nostack']).max() < TOL
    assert np.abs(fe.stack(lp) - G['short_offline']).max() < TOL
    # fewer samples than one window -> no windows, no rows
    tiny = fe.log_power(x[:30])
    assert tiny.shape == (0, 1) and fe.stack(tiny).shape == (0, 5)


@pytest.mark.parametrize('mode', ['exact', 'truncated', 'single'])
def test_chunked_scan_matches_oracle(mode):
    """Time-chunked scan: exact carry (Phi), truncated zero-state pass, and the un-chunked run agree with scipy."""
    sr = 1024
    x = synth.seeg_session(11, 5, sr, 60.0)
    want = O.herff2016_b(x.astype(np.float64), sr, skip_stacking=True)
    fe = FeatureExtractor(sr)
    if mode == 'exact':
        got = fe.log_power(x, chunks=30)                       # 2048-sample chunks << horizon -> Phi carry
        assert fe.scan_plan(len(x), 5, 30)[3] is not None
    elif mode == 'truncated':
        got = fe.log_power(x, chunks=2)                        # 30720-sample chunks > horizon (17408)
        assert fe.scan_plan(len(x), 5, 2)[3] is None and fe.scan_plan(len(x), 5, 2)[2] < 30720
    else:
        got = fe.log_power(x, chunks=1)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < TOL


def test_many_sessions_ragged_channels():
    """Session batch with a channel count that is not a multiple of the warp size."""
    sr = 2048
    xs = np.stack([synth.seeg_session(20 + s, 37, sr, 3.0) for s in range(3)])
    fe = FeatureExtractor(sr)
    got = fe.log_power(xs)
    for s in range(3):
        want = O.herff2016_b(xs[s].astype(np.float64), sr, skip_stacking=True)
        assert np.abs(got[s] - want).max() < TOL


def test_device_resident_tensors():
    import torch
    sr = 1024
    x = synth.seeg_session(12, 16, sr, 5.0)
    fe = FeatureExtractor(sr)
    xd = torch.from_numpy(x).cuda()
    lp = fe.log_power(xd)
    assert lp.is_cuda
    st = fe.stack(lp)
    want = O.herff2016_b(x.astype(np.float64), sr)
    assert np.abs(st.cpu().numpy() - want).max() < TOL


@pytest.mark.parametrize('sr,ln,seconds,n_sess,pieces,check', [
    (1024, 50, 260.0, 3, 4, 1), (2048, 60, 130.0, 3, 4, 1),
    # 3 stream groups x 64 000 samples in 5 pieces: the cut at 76 800 starts group 1 at t = 12 800, inside the tail's horizon
    # (14 912), where the reference's initial state still counts (kappa, sgs/modal.py); session 3 lies in that group
    (1024, 50, 62.5, 6, 5, 3)])
def test_modal_tail_equals_the_zero_state_warm_up(sr, ln, seconds, n_sess, pieces, check, monkeypatch):
    """A piece's start state from the modal sums of the far past + a short cascade run (k_iir_tail, sgs/modal.py) against the
    full-length zero-state warm-up of the same pieces, and against the oracle."""
    import torch
    from sgs import _lib
    n_ch = 12
    xs = np.stack([synth.seeg_session(90 + s, n_ch, sr, seconds) for s in range(n_sess)])
    xd = torch.from_numpy(xs).cuda()
    monkeypatch.setenv('SGS_FEAT_PIECES', '1')
    monkeypatch.setenv('SGS_FEAT_PIECES_P', str(pieces))
    out, launches = {}, {}
    for tail in ('0', '1'):
        monkeypatch.setenv('SGS_FEAT_TAIL', tail)
        fe = FeatureExtractor(sr, line_noise=ln)
        fe.log_power(xd, chunks=3)                                       # plan + tables
        n0 = _lib.launch_count()
        out[tail] = fe.log_power(xd, chunks=3).cpu().numpy()
        launches[tail] = _lib.launch_count() - n0
        if tail == '1':
            t = fe.tail()
            assert t is not None and t.horizon < len(xs[0]) // 2 and t.near_len < fe.horizon() // 4
    assert launches['1'] == launches['0'] + 1, launches                  # the tail kernel ran
    err = np.abs(out['1'] - out['0']).max()
    print('modal tail vs zero-state warm-up: max |diff| %.3g in log-power' % err)
    assert err < 2e-11
    if pieces == 5:
        # a float64 recording goes through the same kernels at full precision (the synthetic data is exact in float32)
        got64 = fe.log_power(xd.double(), chunks=3).cpu().numpy()
        assert np.array_equal(got64, out['1'])
    want = O.herff2016_b(xs[check].astype(np.float64), sr, skip_stacking=True, line_noise=ln)
    assert np.abs(out['1'][check] - want).max() < TOL


@pytest.mark.parametrize('online', [False, True])
def test_balanced_pieces_equal_chunked_scan(online, monkeypatch):
    """The piece decomposition (time lines of all stream groups laid end to end, cut into equal pieces, one or two
    segments per CTA) gives the features of the (group x chunk) decomposition and of the oracle."""
    import torch
    sr, n_ch, seconds, n_sess = 1024, 12, 150.0, 5                       # 60 streams = 2 groups, the second one partial
    xs = np.stack([synth.seeg_session(60 + s, n_ch, sr, seconds) for s in range(n_sess)])
    fe = FeatureExtractor(sr)
    xd = torch.from_numpy(xs).cuda()
    monkeypatch.setenv('SGS_FEAT_PIECES', '0')
    ref = fe.log_power(xd, online=online, chunks=3).cpu().numpy()
    monkeypatch.setenv('SGS_FEAT_PIECES', '1')
    for pieces in ('5', '4', '3'):                                       # pieces straddling the group boundary or not
        monkeypatch.setenv('SGS_FEAT_PIECES_P', pieces)
        got = fe.log_power(xd, online=online, chunks=3).cpu().numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-10                            # prefix-sum re-basing points differ between the two
    if online:
        want = O.ecog_feat_calc(xs[3].astype(np.float64), sr, 50, 10, 4, 5, 50, 32)         # stacked rows: tap 4 = the frame itself
        assert np.abs(got[3] - want[:, 4::5]).max() < 1e-9
    else:
        want = O.herff2016_b(xs[3].astype(np.float64), sr, skip_stacking=True)
        assert np.abs(got[3] - want).max() < 1e-9
