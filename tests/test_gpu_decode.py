"""CUDA LDA decoding, dequantisation and node-semantics Griffin-Lim against the oracle and reference fixtures."""
import os
import pickle

import numpy as np
import pytest

import oracle as O
from sgs import synth
from sgs.features import FeatureExtractor
from sgs.lda import LdaDecoder, pack_estimators
from sgs.griffinlim import GriffinLimNodeOp
from helpers import load, node_noise, GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def model():
    G = load('train_decode.npz')
    with open(os.path.join(GOLDEN, 'estimators.pkl'), 'rb') as fh:
        est = pickle.load(fh)
    return G, est


def test_lda_labels_and_spectrum_match_reference_fixture(model):
    """Class indices are bit-exact against the reference's sklearn predict on its own streamed features."""
    G, est = model
    dec = LdaDecoder(est, G['select'], G['medians'])
    labels, spec = dec.decode(G['dec_feat'], smooth=True)          # already stacked rows (300 x 150)
    assert np.array_equal(labels, G['dec_labels'])
    assert np.array_equal(spec, G['dec_spec'])                      # lookup + scipy-ordered 5-tap smoothing: bit-exact
    labels_b, spec_b = dec.decode(G['dec_feat'], smooth=False)
    assert np.array_equal(spec_b, O.dequantize_spectrogram(labels_b, G['medians']))


def test_lda_fused_gather_from_unstacked_features(model):
    """Decoding straight from the un-stacked log-power array equals stack -> select -> predict (offline and online views)."""
    G, est = model
    sr, bad = int(G['sr']), list(G['bad'])
    x = np.delete(synth.seeg_session(2, int(G['n_ch']), sr, 3.0), bad, axis=1)
    fe = FeatureExtractor(sr)
    dec = LdaDecoder(est, G['select'], G['medians'])
    lp = fe.log_power(x)
    lab_off, spec_off = dec.decode(lp, order=4, step=5, first_row=20)
    assert np.array_equal(lab_off, G['batch_labels'])
    assert np.array_equal(spec_off, G['batch_spec'])
    lp_on = fe.log_power(x, online=True, chunk_size=32)
    lab_on, spec_on = dec.decode(lp_on, order=4, step=5, first_row=0, smooth=True)
    assert np.array_equal(lab_on, G['dec_labels'])
    assert np.array_equal(spec_on, G['dec_spec'])


def test_lda_missing_and_binary_classes():
    """Bins that never saw some classes (train.py:86-91) and sklearn's two-class special case."""
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    rng = np.random.default_rng(5)
    X = rng.normal(size=(600, 12))
    ys = [rng.integers(0, 9, 600), rng.choice([0, 3, 4, 8], 600), rng.choice([2, 7], 600)]
    for y in ys:
        X[np.arange(600), y % 12] += 1.5
    est = [LinearDiscriminantAnalysis().fit(X, y.astype(float)) for y in ys]
    med = np.tile(np.linspace(-12, -3, 9), (3, 1))
    dec = LdaDecoder(est, np.arange(12), med)
    Xt = rng.normal(size=(257, 12))
    labels, spec = dec.decode(Xt)
    want = O.lda_predict(Xt, est, np.arange(12))
    assert np.array_equal(labels, want)
    assert np.array_equal(spec, O.dequantize_spectrogram(want, med))
    W, b, cls = pack_estimators(est)
    assert np.isinf(b[1]).sum() == 5 and np.isinf(b[2]).sum() == 7


def test_griffinlim_node_against_reference_fixture():
    GL = load('griffinlim.npz')
    lm = GL['node_logmel']
    for norm in (1.0, 10.0):
        op = GriffinLimNodeOp(16, 10, 16000, 40, iterations=8, norm_factor=norm)
        pcm = op.synthesize(lm, node_noise(77, len(lm)))
        want = GL['node_pcm_norm%g' % norm]
        assert pcm.shape == want.shape
        d = np.abs(pcm.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3          # int16 truncation may flip one LSB on a rounding boundary
    op = GriffinLimNodeOp(16, 10, 16000, 40)                   # default 5 iterations
    lm2 = GL['node5_logmel']
    d = np.abs(op.synthesize(lm2, node_noise(78, len(lm2))).astype(int) - GL['node5_pcm'].astype(int))
    assert d.max() <= 1


def test_griffinlim_node_long_stream_against_oracle(model):
    """300 frames (crosses the 159/161-sample hops of quirk Q7 at frame 201) + float comparison of blocks."""
    G, _ = model
    spec = G['dec_spec']
    noise = node_noise(4001, len(spec))
    op = GriffinLimNodeOp(16, 10, 16000, 40, iterations=8, norm_factor=10)
    pcm, flt, blk = op.synthesize(spec, noise, want_filtered=True, want_blocks=True)
    ref = O.GriffinLimNode(16, 10, 16000, 40, 8, norm_factor=10)
    want_pcm, want_flt = ref.synthesize(spec, noise)
    assert pcm.shape == G['dec_audio'].shape
    want_blk = ref.block(spec[9:11], noise[10])
    assert np.abs(blk[10] - want_blk).max() <= 1e-11 * np.abs(want_blk).max()
    # the reference's 'ba'-form low-pass has clustered poles near z = -1 (state gain ~6e5): fp64 round-off differences
    # (FMA vs separate multiply-add) are amplified to ~1e-10 relative
    assert np.abs(flt - want_flt).max() <= 1e-8 * np.abs(want_flt).max()
    d = np.abs(pcm.astype(int) - G['dec_audio'].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_griffinlim_sessions_batch_and_device_noise():
    import torch
    med = synth.default_medians()
    lm = synth.logmel_utterances(3, 25, med, seed=3200)
    noise = np.stack([node_noise(90 + s, 25) for s in range(3)])
    op = GriffinLimNodeOp(16, 10, 16000, 40, iterations=8, norm_factor=10)
    pcm = op.synthesize(lm, noise)
    for s in range(3):
        want, _ = O.GriffinLimNode(16, 10, 16000, 40, 8, norm_factor=10).synthesize(lm[s], noise[s])
        assert np.abs(pcm[s].astype(int) - want.astype(int)).max() <= 1
    # device-resident input + counter-based noise: runs, deterministic per seed, differs across seeds
    a = op.synthesize(torch.from_numpy(lm).cuda(), None, seed=1)
    b = op.synthesize(torch.from_numpy(lm).cuda(), None, seed=1)
    c = op.synthesize(torch.from_numpy(lm).cuda(), None, seed=2)
    assert a.is_cuda and torch.equal(a, b) and not torch.equal(a, c)


def test_lda_tensor_core_path_equals_fp64(model, monkeypatch):
    """>= 4096 frames take the tcgen05 split-TF32 scoring + exact fp64 re-scoring of near-ties: labels must equal the
    pure fp64 kernel's (and sklearn's) everywhere, online and offline stacking, ragged last tile, several sessions."""
    G, est = model
    sr, bad = int(G['sr']), list(G['bad'])
    xs = np.stack([np.delete(synth.seeg_session(30 + s, int(G['n_ch']), sr, 25.3), bad, axis=1) for s in range(2)])
    fe = FeatureExtractor(sr)
    lp = fe.log_power(xs, online=True, chunk_size=32)
    assert lp.shape[0] * lp.shape[1] >= 4096
    dec = LdaDecoder(est, G['select'], G['medians'])
    lab_tc, spec_tc = dec.decode(lp, order=4, step=5, first_row=0, smooth=True)
    rescored = dec.last_rescored()
    monkeypatch.setenv('SGS_LDA_TC', '0')
    lab_64, spec_64 = dec.decode(lp, order=4, step=5, first_row=0, smooth=True)
    monkeypatch.delenv('SGS_LDA_TC')
    assert np.array_equal(lab_tc, lab_64) and np.array_equal(spec_tc, spec_64)
    want = O.lda_predict(fe.stack(lp[1], online=True), est, G['select'])
    assert np.array_equal(lab_tc[1], want)
    assert 0 <= rescored < 0.25 * lab_tc.shape[0] * lab_tc.shape[1], rescored       # the filter has to actually filter
    print('tensor-core LDA: %d of %d (frame, bin) pairs re-scored in fp64' % (rescored, lab_tc.shape[0] * lab_tc.shape[1] * 40))
    # offline view (first_row = 20) on device-resident input
    import torch
    lpo = fe.log_power(torch.from_numpy(xs).cuda())
    lab_o, _ = dec.decode(lpo, order=4, step=5, first_row=20)
    monkeypatch.setenv('SGS_LDA_TC', '0')
    lab_o64, _ = dec.decode(lpo, order=4, step=5, first_row=20)
    assert torch.equal(lab_o, lab_o64)


def test_lda_few_frame_kernel_equals_batch_kernel(model):
    """Up to 256 frames per call run k_lda_rows (one block per frame, the streaming nodes' path), more run
    k_lda_decode: same score arithmetic, so the same labels and spectrum to the bit whichever runs."""
    G, est = model
    dec = LdaDecoder(est, G['select'], G['medians'])
    feat = G['dec_feat']                                           # 300 stacked rows -> k_lda_decode
    labels, spec = dec.decode(feat, smooth=True)
    for lo, hi in ((0, 1), (1, 4), (4, 104), (104, 300)):          # -> k_lda_rows
        la, sp = dec.decode(np.ascontiguousarray(feat[lo:hi]), smooth=True)
        assert np.array_equal(la, labels[lo:hi]) and np.array_equal(sp, spec[lo:hi])
    sr, bad = int(G['sr']), list(G['bad'])
    x = np.delete(synth.seeg_session(2, int(G['n_ch']), sr, 3.0), bad, axis=1)
    lp_on = FeatureExtractor(sr).log_power(x, online=True, chunk_size=32)
    la, sp = dec.decode(lp_on, order=4, step=5, first_row=0, n_rows=40, smooth=True)     # gather incl. rows before the stream start
    assert np.array_equal(la, G['dec_labels'][:40]) and np.array_equal(sp, G['dec_spec'][:40])


@pytest.mark.parametrize('groups', [[3] * 20, [7, 6, 6, 6] * 2 + [10], [16, 16, 16, 12], [1, 2, 16, 1, 13, 16, 11]])
def test_gl_node_push_many_frames_equals_one_at_a_time(groups):
    """sgs_gl_node_push with n new frames (all their blocks are synthesised before the first hop is emitted) gives the
    audio of n single-frame pushes: the block ring must outlive a whole push."""
    from sgs import _lib
    rng = np.random.default_rng(0)
    T = sum(groups)
    spec = rng.normal(-2, 1.5, size=(T, 40))
    noise = rng.random((T, 480))

    def run(gs):
        op = GriffinLimNodeOp(16, 10, 16000, 40, 8, 7900, 10)
        pos_all = op.positions(T)
        out, k, prev = [], 0, 0
        pcm = np.empty(16 * 192, np.int16)
        for n in gs:
            n_pcm = _lib.c_int(0)
            fr, nz, ps = (np.ascontiguousarray(a[k:k + n]) for a in (spec, noise, pos_all))
            _lib.check(_lib.lib().sgs_gl_node_push(op.handle(), _lib.ptr(fr), n, _lib.ptr(ps), int(prev), _lib.ptr(nz), 0,
                                                   _lib.ptr(pcm), _lib.C.byref(n_pcm), None))
            out.append(pcm[:n_pcm.value].copy())
            prev, k = int(pos_all[k + n - 1]), k + n
        return np.hstack(out)

    assert np.array_equal(run(groups), run([1] * T))


def test_exp_angle_matches_numpy_and_mpmath():
    """The node's exp(angle(X)) (GriffinLim.py:93) is evaluated by polynomial pieces (csrc/exp_angle.cuh): within a few ulp
    of numpy everywhere, branch cut and signed zeros included, and within 3 ulp of the true value."""
    import mpmath as mp
    from sgs import _lib
    _lib.ensure_init()
    rng = np.random.default_rng(1)
    n = 200000
    re = rng.normal(size=n) * 10.0 ** rng.uniform(-12, 6, n)
    im = rng.normal(size=n) * 10.0 ** rng.uniform(-12, 6, n)
    edge = [(0.0, 1.0), (-0.0, 1.0), (0.0, -1.0), (-0.0, -1.0), (1.0, 0.0), (-1.0, 0.0), (1.0, -0.0), (-1.0, -0.0), (0.0, 0.0),
            (-0.0, 0.0), (0.0, -0.0), (-0.0, -0.0), (1.0, 1.0), (-1.0, 1.0), (1.0, -1.0), (-1.0, -1.0), (1e-300, -1.0),
            (-1e-300, -1.0), (5e-324, -2.0), (-5e-324, -2.0), (3.0, 1e-310), (1e-310, 1e-310), (1e300, -1e300), (1e-16, -1.0)]
    im[:len(edge)] = [e[0] for e in edge]
    re[:len(edge)] = [e[1] for e in edge]
    out = np.empty(n)
    _lib.check(_lib.lib().sgs_exp_angle(_lib.ptr(im), _lib.ptr(re), n, _lib.ptr(out), None))
    want = np.exp(np.arctan2(im, re))          # np.angle(z) = arctan2(z.imag, z.real); building z here would lose the -0.0s
    rel = np.abs(out - want) / want
    assert rel.max() < 2e-15, (rel.max(), im[rel.argmax()], re[rel.argmax()])
    # sides of the branch cut: exp(+pi) vs exp(-pi), a factor 535 apart
    assert np.array_equal(out[:len(edge)] > 20, want[:len(edge)] > 20) and np.array_equal(out[:len(edge)] < 0.05, want[:len(edge)] < 0.05)
    mp.mp.dps = 40
    worst = 0.0
    for i in list(range(len(edge))) + list(range(len(edge), n, 97)):
        if re[i] == 0 or im[i] == 0:
            continue                                             # mpmath has no signed zeros (checked against arctan2 above)
        t = mp.exp(mp.atan2(mp.mpf(float(im[i])), mp.mpf(float(re[i]))))
        worst = max(worst, float(abs(mp.mpf(float(out[i])) - t) / t))
    assert worst < 3 * 2.2e-16, worst


def test_config1_full_size_decode_matches_oracle():
    """BASELINE config 1 at full size - 64 ch @ 1024 Hz, 5 minutes (307 200 samples, 30 000 frames) - decoded on the device
    and by the CPU oracle from the same synthetic recording, random-weight model of the trained architecture and the
    same np.random.rand start of every Griffin-Lim block: features to 1e-9, class indices and spectrogram bit-exact,
    int16 audio within 1 LSB."""
    import decode
    from sgs.synth import default_medians
    sr, n_ch, seconds = 1024, 64, 300.0
    rng = np.random.default_rng(11)
    W = rng.normal(0, 0.3, (40, 9, 150))
    b = rng.normal(0, 1.0, (40, 9))
    cls = np.tile(np.arange(9, dtype=np.float64), (40, 1))
    select = rng.permutation(5 * n_ch)[:150].astype(np.int32)
    medians = default_medians(40, 9)
    x = synth.seeg_session(21, n_ch, sr, seconds)                              # float32 (T x C)
    feats = O.ecog_feat_calc(x.astype(np.float64), sr, 50, 10, 4, 5, 50, 32)
    assert feats.shape == (30000, 5 * n_ch)
    want_labels, _ = O.lda_predict_packed(feats, W, b, cls, select)
    want_spec = O.dequantization_node(want_labels, medians)
    noise = np.random.RandomState(4100).rand(len(want_spec), 480)
    want_pcm, _ = O.GriffinLimNode(16, 10, 16000, 40, 8, norm_factor=10).synthesize(want_spec, noise)

    dec = decode.OfflineDecoder((W, b, cls), medians, select, sr, gl_norm=10, packet_size=32)
    lp = dec.features.log_power(x, online=True, chunk_size=32)
    assert np.abs(dec.features.stack(lp, online=True) - feats).max() < 1e-9
    labels, spec = dec.lda.decode(lp, order=4, step=5, first_row=0, smooth=True)     # 30 000 frames: the tensor-core path
    assert np.array_equal(labels, want_labels)
    assert np.array_equal(spec, want_spec)
    pcm = dec.gl.synthesize(spec, noise)
    assert pcm.shape == want_pcm.shape == (4799840,)
    d = np.abs(pcm.astype(int) - want_pcm.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3


def test_config5_full_length_sessions_causality_and_independence():
    """BASELINE config 5 session shape at full length (128 ch x 600 s @ 2048 Hz): properties that do not need a 10-minute CPU run.
    (1) causality: every stage is causal, so the first frames of the full decode equal the CPU oracle's decode of the first
        seconds alone; (2) independence: a session decodes to the same bits alone or inside a batch."""
    import decode
    import torch
    from sgs.synth import default_medians
    sr, n_ch, seconds, head = 2048, 128, 600.0, 8.0
    rng = np.random.default_rng(13)
    W = rng.normal(0, 0.3, (40, 9, 150)); b = rng.normal(0, 1.0, (40, 9))
    cls = np.tile(np.arange(9, dtype=np.float64), (40, 1))
    select = rng.permutation(5 * n_ch)[:150].astype(np.int32)
    medians = default_medians(40, 9)
    xs = np.stack([synth.seeg_session(40 + s, n_ch, sr, seconds) for s in range(2)])           # (2, 1 228 800, 128) float32
    dec = decode.OfflineDecoder((W, b, cls), medians, select, sr, gl_norm=10, packet_size=64)
    xd = torch.from_numpy(xs).cuda()
    spec2, audio2 = dec.decode(xd, None, seed=3)
    spec2, audio2 = spec2.cpu().numpy(), audio2.cpu().numpy()
    assert spec2.shape == (2, 60000, 40) and audio2.shape == (2, 9599840)
    for s in range(2):
        spec1, audio1 = dec.decode(xd[s:s + 1], None, seed=3)
        assert np.array_equal(spec1[0].cpu().numpy(), spec2[s])
    # audio with the device noise generator depends on (seed, session index): session 0 alone == session 0 in the batch
    spec1, audio1 = dec.decode(xd[0:1], None, seed=3)
    assert np.array_equal(audio1[0].cpu().numpy(), audio2[0])
    # causality against the oracle on the first `head` seconds of session 1
    n_head = int(head * sr)
    feats = O.ecog_feat_calc(xs[1, :n_head].astype(np.float64), sr, 50, 10, 4, 5, 50, 64)
    want_labels, _ = O.lda_predict_packed(feats, W, b, cls, select)
    want_spec = O.dequantization_node(want_labels, medians)
    n = len(want_spec)
    assert n >= 790
    assert np.array_equal(spec2[1, :n], want_spec)
    lp = dec.features.log_power(xd[1], online=True, chunk_size=64)
    assert np.abs(dec.features.stack(lp, online=True)[:n].cpu().numpy() - feats).max() < 1e-9
