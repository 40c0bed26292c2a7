"""train.py on the device against the reference's own training run (tests/golden/train_decode.npz)."""
import pickle

import numpy as np
import pytest

import oracle as O
from sgs import synth, training
from helpers import load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def trained():
    import train
    G = load('train_decode.npz')
    sr, n_ch, dur = int(G['sr']), int(G['n_ch']), float(G['dur'])
    eeg = synth.seeg_session(1, n_ch, sr, dur).astype(np.float64)
    audio = synth.audio_session(1, dur)
    return G, train.train(eeg, audio, sr, 16000, list(G['bad']))


def test_train_matches_reference(trained):
    G, (x_train, q, medians, estimators, select) = trained
    # the device log-mel target agrees with numpy's to ~1e-14 (different FFT operation order), so min/max - and with
    # them the logistic medians and borders - agree to that level; a label can only differ for a value within that
    # distance of a border (documented near-tie)
    assert np.abs(medians - G['medians']).max() < 1e-12
    assert (q[:400] != G['q_head']).sum() <= 1
    hist = np.array([[np.sum(q[:, b] == k) for k in range(9)] for b in range(40)])
    assert np.abs(hist - G['q_hist']).sum() <= 4
    assert np.array_equal(select, G['select'])                       # same 150 features in the same |rho| order
    assert list(x_train.shape) == list(G['x_train_shape'])
    assert np.abs(x_train[:16] - G['x_train_head']).max() < 1e-9
    assert len(estimators) == 40
    for i, e in enumerate(estimators):
        k = len(e.classes_)
        assert k == int(G['n_classes'][i])
        assert np.array_equal(e.classes_, G['classes'][i, :k])
        ref_c, ref_i = G['coef'][i, :e.coef_.shape[0]], G['intercept'][i, :e.intercept_.shape[0]]
        assert np.abs(e.coef_ - ref_c).max() <= 1e-6 * np.abs(ref_c).max()       # eigh-of-Gram vs SVD-of-data
        assert np.abs(e.intercept_ - ref_i).max() <= 1e-6 * max(1.0, np.abs(ref_i).max())
    pickle.loads(pickle.dumps(estimators))                           # the model must stay picklable (train.py:179)


def test_trained_model_predicts_like_reference(trained):
    G, (_, _, medians, estimators, select) = trained
    feats = G['dec_feat']
    got = O.lda_predict(feats, estimators, select)
    assert np.array_equal(got, G['dec_labels'])                      # 300 frames x 40 bins: no label flips


def test_spearman_matches_scipy(trained):
    G, _ = trained
    sr, n_ch, dur, bad = int(G['sr']), int(G['n_ch']), float(G['dur']), list(G['bad'])
    from local.offline import herff2016_b, compute_spectrogram
    x = herff2016_b(np.delete(synth.seeg_session(1, n_ch, sr, dur).astype(np.float64), bad, axis=1), sr)
    y = compute_spectrogram(synth.audio_session(1, dur), 16000, 0.016, 0.01)[20:-4]
    rho, _ = training.spearman(x, y)
    assert np.abs(rho - G['rho']).max() < 1e-12
    # ties: quantised columns have many equal values -> average ranks
    xq = np.round(x[:, :7] * 2) / 2
    from scipy.stats import spearmanr
    tgt = np.mean(y, axis=1)
    want = np.array([spearmanr(xq[:, f], tgt)[0] for f in range(7)])
    got, _ = training.spearman(xq, y)
    assert np.abs(got - want).max() < 1e-12


def test_lda_stats_shards_are_additive():
    rng = np.random.default_rng(3)
    X = rng.normal(9.0, 0.5, (5000, 40))
    lab = rng.integers(0, 9, (5000, 6)).astype(float)
    sel = rng.permutation(40)[:25]
    full = training.lda_stats(X, sel, lab)
    xbar = training.col_means(X, sel)
    assert np.abs(xbar - X[:, sel].mean(0)).max() < 1e-13
    a = training.lda_stats(X[:1800], sel, lab[:1800], xbar=xbar)
    b = training.lda_stats(X[1800:], sel, lab[1800:], xbar=xbar)
    Xc = X[:, sel] - xbar
    assert np.abs(full['G'] - Xc.T @ Xc).max() < 1e-9
    assert np.abs(a['G'] + b['G'] - full['G']).max() < 1e-9
    assert np.array_equal(a['counts'] + b['counts'], full['counts'])
    assert np.abs(a['sums'] + b['sums'] - full['sums']).max() < 1e-9


def _exact_stats(X, sel, lab, xbar, n_classes=9):
    """The statistics the tensor-core path promises, in exact integer arithmetic: Gram matrix and class sums of the
    46-bit fixed-point quantisation of Xc = X[:, sel] - xbar (Python ints never round)."""
    Xc = X[:, sel] - xbar
    amax = np.abs(Xc).max(0)
    e = np.array([np.frexp(a)[1] if a > 0 else 0 for a in amax])
    q = np.rint(Xc * np.ldexp(1.0, 46 - e)).astype(np.int64)
    qo = q.astype(object)
    G = (qo.T @ qo)
    scale = np.ldexp(1.0, e - 46)
    G = np.array([[float(G[i, j]) for j in range(len(sel))] for i in range(len(sel))]) * np.outer(scale, scale)
    nb = lab.shape[1]
    sums = np.zeros((nb, n_classes, len(sel)))
    for b in range(nb):
        for k in range(n_classes):
            rows = lab[:, b] == k
            if rows.any():
                sums[b, k] = np.array([float(v) for v in qo[rows].sum(0)]) * scale
    return G, sums


@pytest.mark.parametrize('n,nf,nb', [(700, 150, 40), (5000, 25, 6), (129, 160, 42), (70000, 33, 3)])
def test_lda_stats_tensor_core_is_exact(n, nf, nb, monkeypatch):
    """tcgen05 kind::i8 digit GEMMs (train_tc.cu) == exact integer Gram / class sums of the quantised data, bit for bit
    up to the final fp64 summation of 11 terms; and within fp64 round-off of the CUDA-core fp64 kernels."""
    rng = np.random.default_rng(n)
    width = max(nf + 10, 48)
    X = rng.normal(9.0, 0.5, (n, width)) * rng.uniform(0.2, 3.0, width)
    X[:, 3] = 7.25                                            # a constant column (zero variance) must not break the scaling
    lab = rng.integers(0, 9, (n, nb)).astype(float)
    lab[:, 0] = rng.choice([2.0, 7.0], n)                     # a bin that never sees most classes
    sel = rng.permutation(width)[:nf]
    tc = training.lda_stats(X, sel, lab)
    monkeypatch.setenv('SGS_TRAIN_TC', '0')
    ref = training.lda_stats(X, sel, lab)
    monkeypatch.delenv('SGS_TRAIN_TC')
    assert np.array_equal(tc['xbar'], ref['xbar']) and np.array_equal(tc['counts'], ref['counts'])
    scale = np.abs(ref['G']).max()
    assert np.abs(tc['G'] - ref['G']).max() <= 1e-11 * scale
    assert np.abs(tc['sums'] - ref['sums']).max() <= 1e-11 * max(1.0, np.abs(ref['sums']).max())
    assert np.array_equal(tc['G'], tc['G'].T)                 # integer accumulation: symmetric to the bit
    if n <= 5000:
        Gx, Sx = _exact_stats(X, sel, lab, tc['xbar'])
        assert np.abs(tc['G'] - Gx).max() <= 4e-16 * scale * 11
        assert np.abs(tc['sums'] - Sx).max() <= 4e-16 * max(1.0, np.abs(Sx).max()) * 6
    # bit-reproducible run to run (integer atomics commute)
    again = training.lda_stats(X, sel, lab)
    assert np.array_equal(again['G'], tc['G']) and np.array_equal(again['sums'], tc['sums'])


@pytest.mark.parametrize('q,n', [(3, 48000 * 7 + 1), (3, 100), (2, 30000), (4, 12345)])
def test_decimate_matches_scipy(q, n):
    """sgs.spectrogram.decimate == scipy.signal.decimate(x, q) (cheby1 order 8, sosfiltfilt, every q-th sample) although
    both passes run as chunked scans with truncated warm-up: agreement to fp64 round-off, edges included."""
    from scipy.signal import decimate as sp_decimate
    from sgs.spectrogram import decimate
    x = synth.audio_session(3, n / 48000.0 + 0.01, 48000)[:n].astype(np.float64)
    x += 0.3                                                  # a DC offset exercises the steady-state initial conditions
    want = sp_decimate(x, q)
    got = decimate(x, q)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def test_cross_validation_driver_above_chance():
    """local/crossval.py (eval_steps/exp1.py's train / decode folds): every held-out stretch is decoded by a model that never
    saw it; on the synthetic session (speech envelope coupled into the high-gamma band) the reconstruction correlates with the
    target far above the alignment-broken control."""
    from local import crossval
    sr, n_ch, seconds = 1024, 16, 60.0
    eeg = synth.seeg_session(5, n_ch, sr, seconds).astype(np.float64)
    audio = synth.audio_session(5, seconds)
    np.random.seed(1)
    reco, orig, wav, (mean, std, rs) = crossval.cross_validate(eeg, audio, sr, 16000, [2], norm_factor=10, nb_folds=3)
    assert reco.shape == orig.shape and reco.shape[1] == 40 and reco.shape[0] > 5900
    assert wav.dtype == np.int16 and len(wav) > 0.98 * 16000 * seconds
    assert len(rs) == 40 and np.isfinite(mean)
    _, _, _, (mean_rand, _, _) = crossval.cross_validate(eeg, audio, sr, 16000, [2], norm_factor=10, nb_folds=3, randomize=True,
                                                         rng=np.random.default_rng(3))
    assert mean > 0.1 and mean > mean_rand + 0.1, (mean, mean_rand)       # measured: 0.21 against -0.02
    assert crossval.last_profile['folds_on_this_rank'] == 3 and crossval.last_profile['total_s'] > 0
    # the device-resident folds (recording uploaded once, model handed from train to decode in the process) give what the
    # reference-signature worker gives from host arrays fold by fold (eval_steps/exp1.py:26-38)
    host = [crossval.train_decode_worker(*a) for a in crossval.construct_folds(eeg, audio, sr, 16000, [2], 10, nb_folds=3)]
    reco_h = np.vstack([r[:min(len(r), len(o))] for _, r, o, _ in host])
    orig_h = np.vstack([np.asarray(o)[:min(len(r), len(o))] for _, r, o, _ in host])
    assert np.array_equal(reco_h, reco) and np.array_equal(orig_h, orig)


def test_hostio_pipelined_transfers_round_trip():
    """sgs.hostio: chunked, multi-threaded pinned staging of pageable arrays - contiguous, a strided channel block, a long
    vector with a ragged tail, a cast, and the small-array fall-back all arrive bit-identical."""
    import torch
    from sgs import hostio
    rng = np.random.default_rng(3)
    big = rng.standard_normal((700000, 48)).astype(np.float32)               # 134 MB
    for a in (big, big[:, 5:21], big[::2, 7:8], rng.standard_normal(40_000_123), rng.standard_normal((50, 3))):
        t = hostio.upload(a)
        assert t.is_cuda and tuple(t.shape) == a.shape
        assert np.array_equal(t.cpu().numpy(), a)
        assert np.array_equal(hostio.download(t), a)
    t = hostio.upload(big[:, :8], np.float64)
    assert t.dtype == torch.float64 and np.array_equal(t.cpu().numpy(), big[:, :8].astype(np.float64))
