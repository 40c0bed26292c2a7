"""Times the UNMODIFIED reference (imported through oracle/ref_shim.py) on this host's CPU cores: the three timings
BASELINE.md section 3 names, each on a bounded sample of the 128-channel @ 2048 Hz workload.

TEST / MEASUREMENT INFRASTRUCTURE: run as a subprocess by bench.py's cpu_baseline leg (rank 0, N = 1) when a copy of the
reference tree is present under the git-ignored baseline/_ref/ (made by __graft_entry__.build() in the build container,
where /root/reference exists; it travels to the GPU box with the snapshot and never enters the history), or by hand:
    SGS_REFERENCE_ROOT=/root/reference python oracle/time_reference.py [chain_seconds] [batch_seconds] [train_seconds]
Prints one JSON line.  The product never imports this file."""
import json
import os
import pickle
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200', 'sgs'))
import ref_shim  # noqa: E402

ref_shim.install()
import synth  # noqa: E402  (sgs/synth.py imported flat: the product's `local` / `livenodes` must not shadow the reference's)
from livenodes import ECogFeatCalc, GriffinLim, LDASynthesis, Dequantization, Node  # noqa: E402
from local.offline import herff2016_b, griffin_lim  # noqa: E402
from local.quantization import dequantize_spectrogram  # noqa: E402
import train as ref_train  # noqa: E402
from sklearn.discriminant_analysis import LinearDiscriminantAnalysis  # noqa: E402


def estimators_from_fixture():
    G = np.load(os.path.join(ROOT, 'tests', 'golden', 'model128.npz'))
    ests = []
    for i in range(40):
        k = int(G['n_classes'][i])
        e = LinearDiscriminantAnalysis()
        e.coef_, e.intercept_, e.classes_ = G['coef'][i, :k].copy(), G['intercept'][i, :k].copy(), G['classes'][i, :k].copy()
        e.n_features_in_ = e.coef_.shape[1]
        ests.append(e)
    return ests, G['select'], G['medians']


def main():
    chain_s = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    batch_s = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
    train_s = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
    sr, n_ch = 2048, 128
    ests, select, medians = estimators_from_fixture()
    out = {"reference_root": ref_shim.REFERENCE_ROOT, "cores": 1, "channels": n_ch, "sample_rate_hz": sr}

    # 1. the as-shipped node chain (decode.py:152-183 wiring), in-process, 64-sample packets
    x = synth.seeg_session(7, n_ch, sr, chain_s).astype(np.float64)
    src = Node.Node(name='src', has_inputs=False)
    fe = ECogFeatCalc.ECogFeatCalc(sr, frame_len_ms=50, frame_shift_ms=10, model_order=4, step_size=5, chunk_size=64)(src)
    lda = LDASynthesis.LDASynthesis(pickle.dumps(ests), select=select)(fe)
    deq = Dequantization.Dequantization(medians)(lda)
    gl = GriffinLim.GriffinLimSynthesis(originalFrameSizeMs=16, frameShiftMs=10, sampleRate=16000, melCoeffCount=40,
                                        numReconstructionIterations=8, normFactor=10)(deq)
    t_in, lat = [0.0], []
    gl.add_output(lambda f: lat.append(time.perf_counter() - t_in[0]))
    np.random.seed(1)
    t0 = time.perf_counter()
    for i in range(0, len(x) - 63, 64):
        t_in[0] = time.perf_counter()
        src.output_data(np.array(x[i:i + 64]))
    dt = time.perf_counter() - t0
    lm = np.array(lat[20:]) * 1e3
    out["as_shipped_chain"] = {"sample": "%g s of one 128-channel session through the reference's livenodes chain, in-process" % chain_s,
                               "seconds": dt, "frames": len(lat), "frames_per_s": len(lat) / dt,
                               "channel_seconds_per_s": n_ch * chain_s / dt,
                               "frame_latency_ms": {"p50": float(np.percentile(lm, 50)), "p99": float(np.percentile(lm, 99))}}

    # 2. the function-level batch path: herff2016_b -> 40 x predict -> dequantize_spectrogram -> offline.griffin_lim
    xb = synth.seeg_session(8, n_ch, sr, batch_s).astype(np.float64)
    t0 = time.perf_counter()
    feat = herff2016_b(xb, sr, 0.05, 0.01)
    t1 = time.perf_counter()
    lab = np.array([e.predict(feat[:, select]) for e in ests]).T
    spec = dequantize_spectrogram(lab.astype(int), medians)
    t2 = time.perf_counter()
    np.random.seed(2)
    griffin_lim(spec)
    t3 = time.perf_counter()
    out["function_batch"] = {"sample": "%g s of one 128-channel session: herff2016_b, 40 x predict, dequantize_spectrogram, offline.griffin_lim" % batch_s,
                             "seconds": t3 - t0, "stage_s": {"features": t1 - t0, "lda_dequantise": t2 - t1, "griffin_lim": t3 - t2},
                             "channel_seconds_per_s": n_ch * batch_s / (t3 - t0)}

    # 3. train.train (decimation of the 48 kHz audio included, as train.py:125 does)
    eeg = synth.seeg_session(9, n_ch, sr, train_s).astype(np.float64)
    # as many spectrogram frames as feature windows (offline.py:100 rounds the window count in floating point; train.py crops
    # 20 frames in front and 4 at the end of the spectrogram)
    n_x = int(np.floor((len(eeg) - 0.05 * sr) / (0.01 * sr))) + 1 - 20
    a16 = synth.audio_session(9, train_s + 1.0)[:(n_x + 24) * 160]
    audio48 = np.repeat(a16, 3)
    t0 = time.perf_counter()
    ref_train.train(eeg, audio48, sr, 48000, [])
    dt = time.perf_counter() - t0
    out["train"] = {"sample": "train.train on %g s of one 128-channel session + 48 kHz audio" % train_s, "seconds": dt,
                    "channel_seconds_per_s": n_ch * train_s / dt}
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
