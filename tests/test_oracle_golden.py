"""The oracle (oracle/oracle.py) against fixtures produced by the reference itself (oracle/gen_golden.py).
CPU only.  Offline features, mel tables, quantisation, LDA coefficients, labels and int16 audio are
bit-exact; streaming features agree to 1 ulp (the node sums a window with one axis-0 reduce, the
closed form sums per column)."""
import os
import pickle

import numpy as np
import pytest

import oracle as O
from sgs import synth
from helpers import load, digest, node_noise, batch_noise, model128, GOLDEN


@pytest.mark.parametrize('sr', [1024, 2048])
@pytest.mark.parametrize('ln', [50, 60])
def test_features_offline_and_online(sr, ln):
    G = load('features.npz')
    x = synth.seeg_session(7, 6, sr, 1.37).astype(np.float64)
    key = 'sr%d_ln%d' % (sr, ln)
    assert digest(x) == str(G[key + '_in'])
    assert np.array_equal(O.herff2016_b(x, sr, line_noise=ln), G[key + '_offline'])
    assert np.array_equal(O.herff2016_b(x, sr, line_noise=ln, skip_stacking=True), G[key + '_offline_nostack'])
    for cs, p in ((32, 16), (32, 32), (64, 64), (32, 100)):
        g = G['%s_online_cs%d_p%d' % (key, cs, p)]
        o = O.ecog_feat_calc(x, sr, 50, 10, 4, 5, ln, cs)
        assert o.shape == g.shape
        assert np.abs(o - g).max() <= 4e-15


def test_features_short_input():
    G = load('features.npz')
    x = synth.seeg_session(8, 1, 1024, 0.30).astype(np.float64)
    assert np.array_equal(O.herff2016_b(x, 1024), G['short_offline'])
    assert np.array_equal(O.herff2016_b(x, 1024, skip_stacking=True), G['short_offline_nostack'])


def test_mel_tables():
    M = load('mel.npz')
    for s in (129, 401):
        m = O.MelFilterBank(s, 40, 16000)
        assert np.array_equal(m.melMatrix, M['mel_%d' % s])
        assert np.array_equal(m.melInvMatrix, M['inv_%d' % s])
    assert np.array_equal(O.MelFilterBank(129, 40, 16000).fromLogMels(M['logmel_in']), M['from_129'])
    assert np.array_equal(O.MelFilterBank(401, 40, 16000).fromLogMels(M['logmel_in']), M['from_401'])
    assert np.array_equal(O.MelFilterBank(129, 40, 16000).toLogMels(M['spec_in']), M['to_129'])


def test_griffinlim_node_bit_exact():
    GL = load('griffinlim.npz')
    lm = GL['node_logmel']
    for norm in (1.0, 10.0):
        pcm, _ = O.GriffinLimNode(16, 10, 16000, 40, 8, norm_factor=norm).synthesize(lm, node_noise(77, len(lm)))
        assert np.array_equal(pcm, GL['node_pcm_norm%g' % norm])
    lm2 = GL['node5_logmel']
    pcm, _ = O.GriffinLimNode(16, 10, 16000, 40, 5).synthesize(lm2, node_noise(78, len(lm2)))
    assert np.array_equal(pcm, GL['node5_pcm'])


@pytest.mark.parametrize('T', [12, 40])
def test_griffinlim_batch_bit_exact(T):
    GL = load('griffinlim.npz')
    pcm = O.griffin_lim_offline(GL['batch_logmel_T%d' % T], batch_noise(500 + T, T))
    assert np.array_equal(pcm, GL['batch_pcm_T%d' % T])


@pytest.fixture(scope='module')
def trained():
    G = load('train_decode.npz')
    sr, n_ch, dur = int(G['sr']), int(G['n_ch']), float(G['dur'])
    eeg = synth.seeg_session(1, n_ch, sr, dur).astype(np.float64)
    audio = synth.audio_session(1, dur)
    assert digest(eeg) == str(G['eeg_digest']) and digest(audio) == str(G['audio_digest'])
    return G, O.train(eeg, audio, sr, list(G['bad']))


def test_train_matches_reference(trained):
    G, (x_train, q, med, est, sel) = trained
    assert np.array_equal(sel, G['select'])
    assert np.array_equal(med, G['medians'])
    assert np.array_equal(q[:400], G['q_head'])
    assert np.array_equal(x_train[:16], G['x_train_head'])
    assert list(x_train.shape) == list(G['x_train_shape'])
    hist = np.array([[np.sum(q[:, b] == k) for k in range(9)] for b in range(40)])
    assert np.array_equal(hist, G['q_hist'])
    for i, e in enumerate(est):
        assert np.array_equal(e.coef_, G['coef'][i, :e.coef_.shape[0]])
        assert np.array_equal(e.intercept_, G['intercept'][i, :e.intercept_.shape[0]])
        assert np.array_equal(e.classes_, G['classes'][i, :len(e.classes_)])
    # the fixture exercises bins with fewer than 9 classes (train.py:86-91)
    assert G['n_classes'].min() < 9


def test_streaming_decode_matches_reference(trained):
    G, (_, _, med, est, sel) = trained
    sr, n_ch, bad = int(G['sr']), int(G['n_ch']), list(G['bad'])
    test = synth.seeg_session(2, n_ch, sr, 3.0).astype(np.float64)
    assert digest(test) == str(G['test_digest'])
    tc = np.delete(test, bad, axis=1)
    nf = len(G['dec_spec'])
    x, lab, spec, pcm, _ = O.decode_streaming(tc, sr, est, sel, med, node_noise(4001, nf), gl_norm=10, chunk_size=32)
    assert np.abs(x - G['dec_feat']).max() <= 4e-15
    assert np.array_equal(lab, G['dec_labels'])
    assert np.array_equal(spec, G['dec_spec'])
    assert np.array_equal(pcm, G['dec_audio'])          # 299 hops incl. the 159/161-sample hops of quirk Q7
    # packed closed form R2 == sklearn predict
    W, b, cls, _ = O.pack_estimators(est)
    lab2, _ = O.lda_predict_packed(x, W, b, cls, sel)
    assert np.array_equal(lab2, lab)
    # the real fork + Manager path (decode.perform_offline_decoding) on the first 1024 samples
    D = load('offline_decoding.npz')
    nf = len(D['spec'])
    _, _, spec, pcm, _ = O.decode_streaming(tc[:1024], sr, est, sel, med, node_noise(4003, nf), gl_norm=10)
    assert np.array_equal(spec, D['spec']) and np.array_equal(pcm, D['audio'])


def test_batch_decode_matches_reference(trained):
    G, (_, _, med, est, sel) = trained
    sr, n_ch, bad = int(G['sr']), int(G['n_ch']), list(G['bad'])
    tc = np.delete(synth.seeg_session(2, n_ch, sr, 3.0).astype(np.float64), bad, axis=1)
    T = len(G['batch_spec'])
    lab, spec, pcm = O.decode_offline_batch(tc, sr, est, sel, med, batch_noise(4002, T))
    assert np.array_equal(lab, G['batch_labels'])
    assert np.array_equal(spec, G['batch_spec'])
    assert np.array_equal(pcm, G['batch_audio'])


def test_pickled_reference_estimators_load():
    with open(os.path.join(GOLDEN, 'estimators.pkl'), 'rb') as fh:
        est = pickle.load(fh)
    G = load('train_decode.npz')
    assert len(est) == 40
    assert np.array_equal(est[5].coef_, G['coef'][5, :est[5].coef_.shape[0]])


def test_model128_streaming_decode_matches_reference():
    """The 128-channel @ 2048 Hz model (reference train.train, 120 s) and the reference node chain's decode of a held-out 2 s
    in 64-sample packets: the oracle's closed form reproduces labels, spectrogram and int16 audio bit for bit from the
    packed coefficients - the model bench.py, smoke() and the full-size GPU tests decode with."""
    (W, b, cls), select, medians, G = model128()
    sr, n_ch = int(G['sr']), int(G['n_ch'])
    assert (sr, n_ch) == (2048, 128) and W.shape == (40, 9, 150)
    test = synth.seeg_session(int(G['test_session']), n_ch, sr, 2.0).astype(np.float64)
    assert digest(test) == str(G['test_digest'])
    x = O.ecog_feat_calc(test, sr, 50, 10, 4, 5, 50, 64)
    lab, _ = O.lda_predict_packed(x, W, b, cls, select)
    assert np.array_equal(lab, G['dec_labels'])
    spec = O.dequantization_node(lab, medians)
    assert np.array_equal(spec, G['dec_spec'])
    pcm, _ = O.GriffinLimNode(16, 10, 16000, 40, 8, norm_factor=10).synthesize(spec, node_noise(int(G['dec_noise_seed']), len(spec)))
    assert np.array_equal(pcm, G['dec_audio'])
    # a trained model: not one class everywhere, and margins much tighter than random weights give
    assert len(np.unique(lab)) >= 5 and G['n_classes'].max() == 9


def test_constructor_variants_match_reference():
    """Options no entry point of the reference passes, pinned against the unmodified nodes (oracle/gen_golden.py:gen_variants):
    ECogFeatCalc(warm_start=False) - cold last filter, no zero fill, the stack buffer starts empty - and
    GriffinLimSynthesis(useLogMels=False) - linear mel input through fromMels."""
    G = load('variants.npz')
    sr, n_ch = int(G['cold_sr']), int(G['cold_n_ch'])
    x = synth.seeg_session(int(G['cold_session']), n_ch, sr, float(G['cold_seconds'])).astype(np.float64)
    for ln in (50, 60):
        want = G['cold_rows_ln%d' % ln]
        got = O.ecog_feat_calc(x, sr, 50, 10, 4, 5, ln, 32, warm_start=False)
        assert got.shape == want.shape and len(want) > 100
        assert np.abs(got - want).max() <= 4e-15
    lin = G['linmel_in']
    node = O.GriffinLimNode(16, 10, 16000, 40, 8, 7900, 10, use_log_mels=False)
    pcm, _ = node.synthesize(lin, node_noise(int(G['linmel_seed']), len(lin)))
    assert np.array_equal(pcm, G['linmel_pcm'])
