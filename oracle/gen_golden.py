"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this container.

Run once here (the reference tree does not exist on the GPU box):
    python oracle/gen_golden.py
The committed fixtures pin oracle/oracle.py (tests/test_oracle_golden.py) and, through it, the
CUDA path (tests/test_gpu_*.py).  Inputs are regenerated from seeds by sgs/synth.py; each fixture
stores a checksum of its inputs so a drifting generator is detected rather than silently accepted.
"""
import hashlib
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200', 'sgs'))

import ref_shim  # noqa: E402

ref_shim.install()

import synth  # noqa: E402  (sgs/synth.py, imported flat so that the product's `local`/`livenodes` do not shadow the reference's)
from livenodes import ECogFeatCalc, GriffinLim, LDASynthesis, Dequantization, Sender, Node  # noqa: E402
from local.offline import herff2016_b, griffin_lim, compute_spectrogram  # noqa: E402
from local.quantization import compute_borders_logistic, quantize_spectrogram, dequantize_spectrogram  # noqa: E402
import local.MelFilterBank as mel  # noqa: E402
import train as ref_train  # noqa: E402
import decode as ref_decode  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


class Collect(Node.Node):
    def __init__(self):
        super().__init__(name='collect', has_outputs=False)
        self.rows = []

    def add_data(self, frame, data_id=0):
        self.rows.append(np.array(frame, copy=True))


def push(node, data, chunk):
    for i in range(0, len(data), chunk):
        node.add_data(np.array(data[i:i + chunk]))


def gen_features():
    out = {}
    for sr in (1024, 2048):
        for ln in (50, 60):
            x = synth.seeg_session(7, 6, sr, 1.37).astype(np.float64)     # 1.37 s: not a multiple of any chunk
            key = 'sr%d_ln%d' % (sr, ln)
            out[key + '_in'] = digest(x)
            out[key + '_offline'] = herff2016_b(x, sr, line_noise=ln)
            out[key + '_offline_nostack'] = herff2016_b(x, sr, line_noise=ln, skip_stacking=True)
            for chunk_size, push_chunk in ((32, 16), (32, 32), (64, 64), (32, 100)):
                node = ECogFeatCalc.ECogFeatCalc(sr, 50, 10, 4, 5, line_noise=ln, chunk_size=chunk_size, has_inputs=False)
                sink = Collect()(node)
                push(node, x, push_chunk)
                out['%s_online_cs%d_p%d' % (key, chunk_size, push_chunk)] = np.array(sink.rows)
    # single channel, very short inputs (edge cases)
    x = synth.seeg_session(8, 1, 1024, 0.30).astype(np.float64)
    out['short_in'] = digest(x)
    out['short_offline'] = herff2016_b(x, 1024)
    out['short_offline_nostack'] = herff2016_b(x, 1024, skip_stacking=True)
    np.savez_compressed(os.path.join(OUT, 'features.npz'), **out)
    print('features.npz', {k: getattr(v, 'shape', v) for k, v in out.items()})


def gen_mel():
    out = {}
    for spec_size, sr in ((129, 16000), (401, 16000), (129, 16000.0)):
        m = mel.MelFilterBank(spec_size, 40, sr)
        out['mel_%d' % spec_size] = m.melMatrix
        out['inv_%d' % spec_size] = m.melInvMatrix
    rng = np.random.default_rng(11)
    lm = rng.uniform(-14, -3, (5, 40))
    out['logmel_in'] = lm
    out['from_129'] = mel.MelFilterBank(129, 40, 16000).fromLogMels(lm)
    out['from_401'] = mel.MelFilterBank(401, 40, 16000).fromLogMels(lm)
    s = rng.uniform(0, 3, (5, 129))
    out['spec_in'] = s
    out['to_129'] = mel.MelFilterBank(129, 40, 16000).toLogMels(s)
    np.savez_compressed(os.path.join(OUT, 'mel.npz'), **out)
    print('mel.npz ok')


def gen_train_decode():
    """Reference train.train on a 24 s x 32-channel session, then the node chain on a held-out 3 s."""
    sr, n_ch, dur = 1024, 32, 24.0
    eeg = synth.seeg_session(1, n_ch, sr, dur).astype(np.float64)
    audio16 = synth.audio_session(1, dur)
    # train.train decimates 48 kHz -> 16 kHz itself (audio-side prep, out of the accelerated path):
    # feed it audio that is already 16 kHz by neutralising the decimate call for this run
    ref_train.decimate = lambda a, q: a
    bad = [3, 17]
    x_train, q, medians, estimators, select = ref_train.train(eeg, audio16, sr, 48000, bad)
    y_spec = compute_spectrogram(audio16, 16000, 0.016, 0.01)[20:-4]
    med2, borders = compute_borders_logistic(y_spec, 9)
    assert np.array_equal(med2, medians)
    out = dict(sr=sr, n_ch=n_ch, dur=dur, bad=np.array(bad), eeg_digest=digest(eeg), audio_digest=digest(audio16),
               select=select, medians=medians, borders=borders,
               q_head=q[:400].astype(np.int8), q_hist=np.array([[np.sum(q[:, b] == k) for k in range(9)] for b in range(40)]),
               y_spec_head=y_spec[:64], x_train_head=x_train[:16], x_train_shape=np.array(x_train.shape),
               n_classes=np.array([len(e.classes_) for e in estimators]))
    W = np.zeros((40, 9, x_train.shape[1]))
    B = np.full((40, 9), np.nan)
    CL = np.full((40, 9), -1.0)
    for i, e in enumerate(estimators):
        k = len(e.classes_)
        W[i, :e.coef_.shape[0]] = e.coef_
        B[i, :e.intercept_.shape[0]] = e.intercept_
        CL[i, :k] = e.classes_
    out.update(coef=W, intercept=B, classes=CL)
    # Spearman coefficients as the reference computes them
    from scipy.stats import spearmanr
    feats_all = herff2016_b(np.delete(eeg, bad, axis=1), sr)
    tgt = np.mean(y_spec, axis=1)
    n = min(len(feats_all), len(tgt))
    out['rho'] = np.array([spearmanr(feats_all[:, f], tgt)[0] for f in range(feats_all.shape[1])])
    out['feats_len'] = np.array([len(feats_all), len(tgt)])

    # ---- decode: node chain in-process (what Sender.sender_process does, Sender.py:23-36) ----
    test = synth.seeg_session(2, n_ch, sr, 3.0).astype(np.float64)
    out['test_digest'] = digest(test)
    pickled = pickle.dumps(estimators)
    src = Node.Node(name='src', has_inputs=False)
    r_seeg, r_spec, r_audio = [], [], []
    from livenodes import ChannelSelector
    sel = ChannelSelector.ChannelSelector(exclude=bad)(src)
    fe = ECogFeatCalc.ECogFeatCalc(sr, frame_len_ms=50, frame_shift_ms=10, model_order=4, step_size=5, chunk_size=32)(sel)
    lda = LDASynthesis.LDASynthesis(pickled, select=select)(fe)
    deq = Dequantization.Dequantization(medians)(lda)
    gl = GriffinLim.GriffinLimSynthesis(originalFrameSizeMs=16, frameShiftMs=10, sampleRate=16000, melCoeffCount=40,
                                        numReconstructionIterations=8, normFactor=10)(deq)
    r_feat, r_lab = [], []
    fe.add_output(lambda f: r_feat.append(np.array(f, copy=True)))
    lda.add_output(lambda f: r_lab.append(np.array(f, copy=True)))
    deq.add_output(lambda f: r_spec.append(np.array(f, copy=True)))
    gl.add_output(lambda f: r_audio.append(np.array(f, copy=True)))
    np.random.seed(4001)
    for i in range(0, len(test), 16):                                   # decode.py:77 -> 16 ms @1024 Hz = 16 samples
        src.output_data(np.array(test[i:i + 16]))
    out['dec_feat'] = np.array(r_feat)
    out['dec_labels'] = np.array(r_lab).astype(np.int8)
    out['dec_spec'] = np.array(r_spec)
    out['dec_audio'] = np.hstack([a for a in r_audio if len(a)])
    out['dec_noise_seed'] = 4001
    # function-level batch path on the same data (CPU baseline ii)
    xb = herff2016_b(np.delete(test, bad, axis=1), sr)
    lab_b = np.array([e.predict(xb[:, select]) for e in estimators]).T
    spec_b = dequantize_spectrogram(lab_b, medians)
    np.random.seed(4002)
    audio_b = griffin_lim(spec_b)
    out.update(batch_labels=lab_b.astype(np.int8), batch_spec=spec_b, batch_audio=audio_b, batch_noise_seed=4002)
    with open(os.path.join(OUT, 'estimators.pkl'), 'wb') as fh:
        pickle.dump(estimators, fh)
    np.savez_compressed(os.path.join(OUT, 'train_decode.npz'), **out)
    print('train_decode.npz', {k: getattr(v, 'shape', v) for k, v in out.items()})

    # ---- the real fork + Manager path once, tiny (decode.py:71-96) ----
    tiny = test[:1024]
    np.random.seed(4003)
    spec, audio, rec, sf = ref_decode.perform_offline_decoding((pickled, medians, bad, select), tiny, sr, 10)
    np.savez_compressed(os.path.join(OUT, 'offline_decoding.npz'), spec=spec, audio=audio, rec_shape=np.array(rec.shape),
                        noise_seed=4003, n=1024)
    print('offline_decoding.npz', spec.shape, audio.shape, rec.shape)


def gen_model128():
    """The decode model of the 128-channel configurations ("produced by train.train on 120 s of the same
    generator"): the UNMODIFIED reference train.train on a 120 s x 128-channel @ 2048 Hz synthetic session, packed as dense
    arrays (coef_/intercept_/classes_ per mel bin) so bench.py, smoke() and the full-size GPU tests decode with a TRAINED
    model instead of random weights, plus the reference node chain's output on 2 s of a held-out session of the same shape."""
    sr, n_ch, dur = 2048, 128, 120.0
    eeg = synth.seeg_session(101, n_ch, sr, dur).astype(np.float64)
    audio16 = synth.audio_session(101, dur)
    ref_train.decimate = lambda a, q: a
    x_train, q, medians, estimators, select = ref_train.train(eeg, audio16, sr, 48000, [])
    W = np.zeros((40, 9, x_train.shape[1]))
    B = np.full((40, 9), -np.inf)
    CL = np.zeros((40, 9))
    for i, e in enumerate(estimators):
        k = len(e.classes_)
        assert e.coef_.shape[0] == k, 'binary bin in the 128-channel model: pack it by hand'
        W[i, :k], B[i, :k], CL[i, :k] = e.coef_, e.intercept_, e.classes_
    out = dict(sr=sr, n_ch=n_ch, dur=dur, train_session=101, eeg_digest=digest(eeg), audio_digest=digest(audio16),
               select=select.astype(np.int32), medians=medians, coef=W, intercept=B, classes=CL,
               n_classes=np.array([len(e.classes_) for e in estimators]),
               train_accuracy=np.array([np.mean(e.predict(x_train) == q[:len(x_train), i]) for i, e in enumerate(estimators)]))
    # held-out 2 s through the reference node chain (decode.py wiring, 64-sample packets as decode.py:116 sets for 2048 Hz)
    test = synth.seeg_session(102, n_ch, sr, 2.0).astype(np.float64)
    src = Node.Node(name='src', has_inputs=False)
    fe = ECogFeatCalc.ECogFeatCalc(sr, frame_len_ms=50, frame_shift_ms=10, model_order=4, step_size=5, chunk_size=64)(src)
    lda = LDASynthesis.LDASynthesis(pickle.dumps(estimators), select=select)(fe)
    deq = Dequantization.Dequantization(medians)(lda)
    gl = GriffinLim.GriffinLimSynthesis(originalFrameSizeMs=16, frameShiftMs=10, sampleRate=16000, melCoeffCount=40,
                                        numReconstructionIterations=8, normFactor=10)(deq)
    r_lab, r_spec, r_audio = [], [], []
    lda.add_output(lambda f: r_lab.append(np.array(f, copy=True)))
    deq.add_output(lambda f: r_spec.append(np.array(f, copy=True)))
    gl.add_output(lambda f: r_audio.append(np.array(f, copy=True)))
    np.random.seed(4010)
    for i in range(0, len(test), 64):
        src.output_data(np.array(test[i:i + 64]))
    out.update(test_session=102, test_digest=digest(test), dec_labels=np.array(r_lab).astype(np.int8), dec_spec=np.array(r_spec),
               dec_audio=np.hstack([a for a in r_audio if len(a)]), dec_noise_seed=4010)
    np.savez_compressed(os.path.join(OUT, 'model128.npz'), **out)
    print('model128.npz', {k: getattr(v, 'shape', v) for k, v in out.items()})
    print('train accuracy per bin (mean %.3f)' % out['train_accuracy'].mean())


def gen_griffinlim():
    out = {}
    med = synth.default_medians()
    lm = synth.logmel_utterances(1, 14, med, seed=3001)[0]
    out['node_logmel'] = lm
    for norm in (1.0, 10.0):
        node = GriffinLim.GriffinLimSynthesis(16, 10, 16000, 40, numReconstructionIterations=8, normFactor=norm)
        got = []
        node.add_output(lambda f: got.append(np.array(f, copy=True)))
        np.random.seed(77)
        for k in range(len(lm)):
            node.add_data(lm[k])
        out['node_pcm_norm%g' % norm] = np.hstack([g for g in got if len(g)])
    out['node_seed'] = 77
    # numReconstructionIterations default (5) on a different sequence
    lm2 = synth.logmel_utterances(1, 6, med, seed=3002)[0]
    node = GriffinLim.GriffinLimSynthesis(16, 10, 16000, 40)
    got = []
    node.add_output(lambda f: got.append(np.array(f, copy=True)))
    np.random.seed(78)
    for k in range(len(lm2)):
        node.add_data(lm2[k])
    out['node5_logmel'] = lm2
    out['node5_pcm'] = np.hstack([g for g in got if len(g)])
    # batch form
    for T in (12, 40):
        lmb = synth.logmel_utterances(1, T, med, seed=3100 + T)[0]
        np.random.seed(500 + T)
        out['batch_logmel_T%d' % T] = lmb
        out['batch_pcm_T%d' % T] = griffin_lim(lmb)
    np.savez_compressed(os.path.join(OUT, 'griffinlim.npz'), **out)
    print('griffinlim.npz', {k: getattr(v, 'shape', v) for k, v in out.items()})


def gen_variants():
    """Constructor options no entry point of the reference passes, run through the unmodified nodes: ECogFeatCalc with
    warm_start=False (cold last filter, no zero fill, the stack buffer starts empty) and GriffinLimSynthesis with
    useLogMels=False (linear mel input through fromMels)."""
    out = {}
    sr, n_ch = 1024, 6
    x = synth.seeg_session(61, n_ch, sr, 1.5).astype(np.float64)
    for ln in (50, 60):
        src = Node.Node(name='src', has_inputs=False)
        fe = ECogFeatCalc.ECogFeatCalc(sr, frame_len_ms=50, frame_shift_ms=10, model_order=4, step_size=5, line_noise=ln,
                                       warm_start=False, chunk_size=32)(src)
        rows = []
        fe.add_output(lambda f: rows.append(np.array(f, copy=True)))
        for i in range(0, len(x), 32):
            src.output_data(np.array(x[i:i + 32]))
        out['cold_rows_ln%d' % ln] = np.array(rows)
    out['cold_session'], out['cold_sr'], out['cold_n_ch'], out['cold_seconds'] = 61, sr, n_ch, 1.5
    med = synth.default_medians()
    lin = np.exp(synth.logmel_utterances(1, 12, med, seed=3003)[0])
    node = GriffinLim.GriffinLimSynthesis(16, 10, 16000, 40, numReconstructionIterations=8, normFactor=10, useLogMels=False)
    got = []
    node.add_output(lambda f: got.append(np.array(f, copy=True)))
    np.random.seed(79)
    for k in range(len(lin)):
        node.add_data(lin[k])
    out['linmel_in'], out['linmel_pcm'], out['linmel_seed'] = lin, np.hstack([g for g in got if len(g)]), 79
    np.savez_compressed(os.path.join(OUT, 'variants.npz'), **out)
    print('variants.npz', {k: getattr(v, 'shape', v) for k, v in out.items()})


if __name__ == '__main__':
    only = sys.argv[1:]
    for fn in (gen_features, gen_mel, gen_griffinlim, gen_train_decode, gen_model128, gen_variants):
        if not only or fn.__name__ in only:
            fn()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
