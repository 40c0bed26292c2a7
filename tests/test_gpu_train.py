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
