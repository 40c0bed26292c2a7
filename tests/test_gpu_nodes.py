"""The livenodes drop-in nodes (streaming device state) and decode.py entry points against reference fixtures."""
import os
import pickle

import numpy as np
import pytest

from sgs import synth
from helpers import load, GOLDEN

pytestmark = pytest.mark.gpu


def push(node, data, chunk):
    for i in range(0, len(data), chunk):
        node.add_data(np.array(data[i:i + chunk]))


@pytest.mark.parametrize('sr', [1024, 2048])
@pytest.mark.parametrize('ln', [50, 60])
def test_ecog_feat_calc_node_streaming(sr, ln):
    from livenodes import ECogFeatCalc
    G = load('features.npz')
    x = synth.seeg_session(7, 6, sr, 1.37).astype(np.float64)
    for chunk_size, push_chunk in ((32, 16), (32, 32), (64, 64), (32, 100)):
        node = ECogFeatCalc.ECogFeatCalc(sr, 50, 10, 4, 5, line_noise=ln, chunk_size=chunk_size, has_inputs=False)
        rows = []
        node.add_output(lambda f: rows.append(np.array(f, copy=True)))
        push(node, x, push_chunk)
        g = G['sr%d_ln%d_online_cs%d_p%d' % (sr, ln, chunk_size, push_chunk)]
        got = np.array(rows)
        assert got.shape == g.shape
        assert np.abs(got - g).max() < 1e-9


@pytest.fixture(scope='module')
def model():
    G = load('train_decode.npz')
    with open(os.path.join(GOLDEN, 'estimators.pkl'), 'rb') as fh:
        blob = fh.read()
    return G, blob


def test_node_chain_matches_reference_chain(model):
    """decode.setup_decoder graph fed in-process with 16-sample chunks == the reference chain's recorded outputs."""
    import decode
    from livenodes import Node
    G, blob = model
    sr, bad = int(G['sr']), list(G['bad'])
    test = synth.seeg_session(2, int(G['n_ch']), sr, 3.0).astype(np.float64)
    src = Node.Node(name='src', has_inputs=False)
    rec_seeg, rec_spec, rec_audio = decode.setup_decoder(src, sr, blob, G['medians'], bad, G['select'], gl_norm=10,
                                                         include_soundcard=False)
    np.random.seed(4001)
    for i in range(0, len(test), 16):
        src.output_data(np.array(test[i:i + 16]))
    spec = np.array(rec_spec.get_data())
    audio = np.hstack([a for a in rec_audio.get_data() if len(a)])
    assert np.array_equal(spec, G['dec_spec'])
    d = np.abs(audio.astype(int) - G['dec_audio'].astype(int))
    assert audio.shape == G['dec_audio'].shape and d.max() <= 1 and (d > 0).mean() < 1e-3
    assert np.vstack(rec_seeg.get_data()).shape == test.shape


def test_perform_offline_decoding_batched(model):
    import decode
    G, blob = model
    D = load('offline_decoding.npz')
    sr, bad = int(G['sr']), list(G['bad'])
    tiny = synth.seeg_session(2, int(G['n_ch']), sr, 3.0).astype(np.float64)[:1024]
    np.random.seed(4003)
    spec, audio, rec, sf = decode.perform_offline_decoding((blob, G['medians'], bad, G['select']), tiny, sr, 10)
    assert sf == sr and rec.shape == tiny.shape
    assert np.array_equal(spec, D['spec'])
    assert audio.dtype == np.int16 and audio.shape == D['audio'].shape
    assert np.abs(audio.astype(int) - D['audio'].astype(int)).max() <= 1


FORKED = r"""
import os, sys, pickle, numpy as np
root = sys.argv[1]
sys.path.insert(0, os.path.join(root, 'closed-loop-seeg-speech-synthesis_b200'))
from sgs import synth
import decode
G = np.load(os.path.join(root, 'tests', 'golden', 'train_decode.npz'))
blob = open(os.path.join(root, 'tests', 'golden', 'estimators.pkl'), 'rb').read()
sr, bad = int(G['sr']), list(G['bad'])
tiny = synth.seeg_session(2, int(G['n_ch']), sr, 3.0).astype(np.float64)[:1024]
np.random.seed(4003)
# the parent process never touches CUDA: the graph (and its device state) lives in the forked feeder process
spec, audio, rec, sf = decode.perform_offline_decoding((blob, G['medians'], bad, G['select']), tiny, sr, 10, streaming=True)
np.savez(sys.argv[2], spec=spec, audio=audio, rec_shape=np.array(rec.shape))
"""


def test_perform_offline_decoding_forked_sender(model, tmp_path):
    """The reference's own execution model (decode.py:71-96): Sender forks, the whole graph runs in the child,
    results come back through Manager lists.  Needs a parent without a CUDA context, hence the subprocess."""
    import subprocess
    import sys
    D = load('offline_decoding.npz')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / 'forked.npz')
    subprocess.run([sys.executable, '-c', FORKED, root, out], check=True, timeout=600)
    R = np.load(out)
    assert np.array_equal(R['spec'], D['spec'])
    assert R['audio'].shape == D['audio'].shape
    assert np.abs(R['audio'].astype(int) - D['audio'].astype(int)).max() <= 1
    assert list(R['rec_shape']) == list(D['rec_shape'])


def test_dequantize_spectrogram_and_herff():
    from local.offline import herff2016_b
    from local.quantization import dequantize_spectrogram, quantize_spectrogram, compute_borders_logistic
    import oracle as O
    G = load('train_decode.npz')
    q = G['q_head'].astype(float)
    assert np.array_equal(dequantize_spectrogram(q, G['medians']), O.dequantize_spectrogram(q, G['medians']))
    F = load('features.npz')
    x = synth.seeg_session(7, 6, 1024, 1.37).astype(np.float64)
    assert np.abs(herff2016_b(x, 1024) - F['sr1024_ln50_offline']).max() < 1e-9
    assert np.abs(herff2016_b(x, 1024, line_noise=60, skip_stacking=True) - F['sr1024_ln60_offline_nostack']).max() < 1e-9
    y = G['y_spec_head']
    med, borders = compute_borders_logistic(y, 9)
    assert np.array_equal(quantize_spectrogram(y, borders), O.quantize_spectrogram(y, borders))


def test_streamed_host_decode_equals_single_call(model):
    """OfflineDecoder.decode with several host sessions (double-buffered H2D / compute / D2H) == resident decode."""
    import torch
    import decode
    G, blob = model
    sr, bad = int(G['sr']), list(G['bad'])
    xs = np.stack([np.delete(synth.seeg_session(40 + s, int(G['n_ch']), sr, 4.0), bad, axis=1) for s in range(3)])
    dec = decode.OfflineDecoder(blob, G['medians'], G['select'], sr, gl_norm=10, packet_size=32)
    spec_h, audio_h = dec.decode(xs, None, seed=5)
    assert isinstance(spec_h, np.ndarray) and audio_h.dtype == np.int16
    for s in range(3):
        spec_d, audio_d = dec.decode(torch.from_numpy(xs[s:s + 1]).cuda(), None, seed=5 + s)
        assert np.array_equal(spec_h[s], spec_d[0].cpu().numpy())
        assert np.array_equal(audio_h[s], audio_d[0].cpu().numpy())
    spec_p, audio_p = dec.decode(xs, None, seed=5, pinned_outputs=True)
    assert np.array_equal(spec_p, spec_h) and np.array_equal(audio_p, audio_h)


def _run_graph(model, fused, packet, seconds=2.0, chunk_size=32, tap=False):
    import decode
    from livenodes import Node
    G, blob = model
    sr, bad = int(G['sr']), list(G['bad'])
    test = synth.seeg_session(3, int(G['n_ch']), sr, seconds)               # float32 packets, like an amplifier delivers
    os.environ['SGS_FUSED_CHAIN'] = '1' if fused else '0'
    try:
        src = Node.Node(name='src', has_inputs=False)
        rec_seeg, rec_spec, rec_audio = decode.setup_decoder(src, sr, blob, G['medians'], bad, G['select'], gl_norm=10,
                                                             packet_size=chunk_size, include_soundcard=False)
        feat_node = src.output_classes[0].output_classes[0]
        lda_node = feat_node.output_classes[0]
        rows, labels = [], []
        feat_node.add_output(lambda f: rows.append(np.array(f, copy=True)))
        lda_node.add_output(lambda f: labels.append(np.array(f, copy=True)))
        np.random.seed(4001)
        for i in range(0, len(test), packet):
            src.output_data(np.array(test[i:i + packet]))
        assert (feat_node._chain is not None) == fused
    finally:
        os.environ.pop('SGS_FUSED_CHAIN', None)
    audio = [a for a in rec_audio.get_data()]
    return np.array(rows), np.array(labels), np.array(rec_spec.get_data()), audio


@pytest.mark.parametrize('packet,chunk_size', [(16, 32), (64, 64), (100, 32), (700, 32)])
def test_fused_chain_equals_node_by_node(model, packet, chunk_size):
    """One sgs_chain_push per packet delivers exactly what the four nodes deliver one call at a time: same callbacks,
    same order, same bits (features, labels, smoothed spectrum, int16 hops incl. the 159/161-sample ones)."""
    a = _run_graph(model, True, packet, chunk_size=chunk_size)
    b = _run_graph(model, False, packet, chunk_size=chunk_size)
    assert len(a[0]) > 150
    for x, y in zip(a[:3], b[:3]):
        assert x.shape == y.shape and np.array_equal(x, y)
    assert len(a[3]) == len(b[3]) and len(a[3]) == len(a[0]) - 1            # the first frame emits no audio (GriffinLim.py:131)
    assert all(np.array_equal(p, q) for p, q in zip(a[3], b[3]))
    assert {len(p) for p in a[3]} <= {159, 160, 161}


@pytest.mark.parametrize('fused', [True, False])
def test_write_head_rebase_changes_nothing(model, fused, monkeypatch):
    """Write-head positions travel as int32 sample counts (37 h at 16 kHz); the node shifts them down before they outgrow that
    (sgs_gl_node_rebase) - here every 2000 samples, 15 times in 2 s - and emits the same hops as without."""
    from livenodes import GriffinLim
    want = _run_graph(model, fused, 64, chunk_size=64)
    monkeypatch.setattr(GriffinLim.GriffinLimSynthesis, '_REBASE_AT', 2000)
    got = _run_graph(model, fused, 64, chunk_size=64)
    assert len(got[3]) == len(want[3]) > 150 and all(np.array_equal(p, q) for p, q in zip(got[3], want[3]))


@pytest.mark.parametrize('ln', [50, 60])
def test_ecog_feat_calc_cold_start_matches_reference(ln):
    """ECogFeatCalc(warm_start=False) against the unmodified reference node (tests/golden/variants.npz): cold last filter,
    frames from the first sample on, no output before the stack buffer holds 21 rows."""
    from livenodes import Node, ECogFeatCalc
    G = load('variants.npz')
    sr, n_ch = int(G['cold_sr']), int(G['cold_n_ch'])
    x = synth.seeg_session(int(G['cold_session']), n_ch, sr, float(G['cold_seconds'])).astype(np.float64)
    want = G['cold_rows_ln%d' % ln]
    for packet in (32, 100):
        src = Node.Node(name='src', has_inputs=False)
        fe = ECogFeatCalc.ECogFeatCalc(sr, 50, 10, 4, 5, line_noise=ln, warm_start=False, chunk_size=32)(src)
        rows = []
        fe.add_output(lambda f: rows.append(np.array(f, copy=True)))
        for i in range(0, len(x), packet):
            src.output_data(np.array(x[i:i + packet]))
        got = np.array(rows)
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-9


def test_griffinlim_linear_mels_match_reference():
    """GriffinLimSynthesis(useLogMels=False) - linear mel frames through fromMels - against the unmodified reference node."""
    from livenodes import GriffinLim
    G = load('variants.npz')
    lin = G['linmel_in']
    node = GriffinLim.GriffinLimSynthesis(16, 10, 16000, 40, numReconstructionIterations=8, normFactor=10, useLogMels=False)
    got = []
    node.add_output(lambda f: got.append(np.array(f, copy=True)))
    np.random.seed(int(G['linmel_seed']))
    for k in range(len(lin)):
        node.add_data(lin[k])
    pcm = np.hstack([g for g in got if len(g)])
    want = G['linmel_pcm']
    d = np.abs(pcm.astype(int) - want.astype(int))
    assert pcm.shape == want.shape and d.max() <= 1 and (d > 0).mean() < 2e-3


def test_fused_chain_not_used_for_other_wirings(model):
    """A second consumer between the nodes, or a missing node, keeps the per-node path."""
    from livenodes import Node, ECogFeatCalc, LDASynthesis, Dequantization, LambdaNode
    G, blob = model
    sr = int(G['sr'])
    n_good = int(G['n_ch']) - len(list(G['bad']))
    x = synth.seeg_session(3, n_good, sr, 0.5)
    src = Node.Node(name='src', has_inputs=False)
    feat = ECogFeatCalc.ECogFeatCalc(sr, 50, 10, 4, 5, chunk_size=32)(src)
    lda = LDASynthesis.LDASynthesis(blob, select=G['select'])(feat)
    mid = LambdaNode.LambdaNode(lambda f: f)(lda)                          # not the reference wiring
    deq = Dequantization.Dequantization(G['medians'])(mid)
    out = []
    deq.add_output(out.append)
    for i in range(0, len(x), 32):
        src.output_data(np.array(x[i:i + 32]))
    assert feat._chain is None and len(out) > 30


def test_config2_streaming_chain_matches_oracle():
    """BASELINE config 2 shape: 128 ch @ 2048 Hz pushed in 64-sample float32 packets through the decode.setup_decoder graph
    (fused chain) == the CPU oracle's closed form of the reference chain on the same recording, model and noise draws."""
    import decode
    import oracle as O
    from livenodes import Node
    from sgs.synth import default_medians
    sr, n_ch, seconds = 2048, 128, 6.0
    rng = np.random.default_rng(12)
    W = rng.normal(0, 0.3, (40, 9, 150)); b = rng.normal(0, 1.0, (40, 9))
    cls = np.tile(np.arange(9, dtype=np.float64), (40, 1))
    select = rng.permutation(5 * n_ch)[:150].astype(np.int32)
    medians = default_medians(40, 9)

    from sgs.training import PackedLDA
    ests = []
    for i in range(40):
        e = PackedLDA()
        e.coef_, e.intercept_, e.classes_ = W[i], b[i], cls[i]
        ests.append(e)
    blob = pickle.dumps(ests)
    x = synth.seeg_session(31, n_ch, sr, seconds)
    src = Node.Node(name='src', has_inputs=False)
    rec_seeg, rec_spec, rec_audio = decode.setup_decoder(src, sr, blob, medians, [], select, gl_norm=10, packet_size=64,
                                                         include_soundcard=False)
    feat_node = src.output_classes[0].output_classes[0]
    rows = []
    feat_node.add_output(lambda f: rows.append(np.array(f, copy=True)))
    np.random.seed(4200)
    for i in range(0, len(x), 64):
        src.output_data(np.array(x[i:i + 64]))
    assert feat_node._chain is not None
    spec = np.array(rec_spec.get_data())
    audio = np.hstack(rec_audio.get_data())
    n_frames = len(spec)
    noise = np.zeros((n_frames, 480))
    rs = np.random.RandomState(4200)
    for k in range(1, n_frames):
        noise[k] = rs.rand(480)
    wx, wl, ws, wp, _ = O.decode_streaming(x.astype(np.float64), sr, pickle.loads(blob), select, medians, noise, gl_norm=10, chunk_size=64)
    assert n_frames == len(ws) and n_frames > 590
    assert np.abs(np.array(rows) - wx).max() < 1e-9
    assert np.array_equal(spec, ws)
    d = np.abs(audio.astype(int) - wp.astype(int))
    assert audio.shape == wp.shape and d.max() <= 1 and (d > 0).mean() < 2e-3


@pytest.mark.parametrize('warm_start', [False, True])
@pytest.mark.parametrize('chunk', [16, 100, 5000])
def test_filtering_framebuffer_matches_scipy(warm_start, chunk):
    """FrameBuffer(filter_coefficients=sos) (livenodes/FrameBuffer.py:86-177): causal sosfilt with carried state on the device,
    cold start (zi * first sample) or warm start (unit zi, zero fill pushed through the filter first), then framing with the
    reference's rounded fractional shifts.  Checked against scipy.signal.sosfilt on the whole recording."""
    from scipy.signal import iirfilter, sosfilt, sosfilt_zi
    from livenodes import FrameBuffer
    sr, n_ch = 1024, 7
    x = synth.seeg_session(9, n_ch, sr, 6.0).astype(np.float64)
    sos = iirfilter(8, [70 / (sr / 2), 170 / (sr / 2)], btype='band', ftype='butter', output='sos')
    fb = FrameBuffer.FrameBuffer(50, 10, sr, filter_coefficients=sos, warm_start=warm_start)
    frames = []
    fb.add_output(lambda f: frames.append(np.array(f, copy=True)))
    for i in range(0, len(x), chunk):
        fb.add_data(x[i:i + chunk])
    fs = int(0.05 * sr)
    zi = np.repeat(sosfilt_zi(sos)[:, :, None], n_ch, axis=2)
    if warm_start:
        fill = fs - int(0.01 * sr)
        y, _ = sosfilt(sos, np.vstack([np.zeros((fill, n_ch)), x]), axis=0, zi=zi)
    else:
        y, _ = sosfilt(sos, x, axis=0, zi=zi * x[0])
    first_ms = fs / float(sr) * 1000.0
    want, k, e = [], 0, fs
    while e <= len(y):
        want.append(y[e - fs:e])
        k += 1
        e = round(((first_ms + k * 10.0) / 1000.0) * float(sr))
    assert len(frames) == len(want) and len(frames) > 550
    got, want = np.array(frames), np.array(want)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
