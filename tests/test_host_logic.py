"""Host-side tables and runtime on CPU: window schedules, mel tables, scan planning, the Node callback runtime."""
import os
import time

import numpy as np
import pytest

import oracle as O
from sgs import design
from sgs.features import FeatureExtractor
from helpers import load


@pytest.mark.parametrize('sr', [1024, 2048])
def test_window_tables_match_the_oracle_schedules(sr):
    p = design.FeaturePlan(sr)
    n = int(7.3 * sr)
    starts, wl = p.offline_window_starts(n)
    assert wl == int(round(0.05 * sr)) and len(starts) == int(np.floor((n - 0.05 * sr) / (0.01 * sr))) + 1
    assert all(starts[k] == int(round(k * 0.01 * sr)) for k in (0, 1, 23, 48, len(starts) - 1))
    # online schedule: frame ends of the closed form used by the oracle (FrameBuffer.py:177)
    usable = (n // 32) * 32
    ends = p.online_frame_ends(usable)
    frame_size = int(0.05 * sr)
    first_ms = frame_size / float(sr) * 1000.0
    k, e, want = 0, frame_size, []
    while e <= usable + p.zero_fill:
        want.append(e); k += 1
        e = round(((first_ms + k * 10.0) / 1000.0) * float(sr))
    assert list(ends) == want
    assert p.frame_size == {1024: 51, 2048: 102}[sr] and p.zero_fill == {1024: 41, 2048: 82}[sr]


def test_mel_tables_match_reference_fixture():
    M = load('mel.npz')
    for size in (129, 401):
        t = design.MelTables(size, 40, 16000)
        assert np.array_equal(t.mel, M['mel_%d' % size]) and np.array_equal(t.inv, M['inv_%d' % size])
        dense = np.zeros_like(t.inv)
        for f in range(size):
            for j in range(2):
                if t.inv_w[f, j] != 0:
                    dense[t.inv_idx[f, j], f] = t.inv_w[f, j]
        assert np.array_equal(dense, t.inv)                  # the 2-tap form loses nothing


def test_known_constants():
    assert np.allclose(design.gaussian_taps(0.5), [2.63865e-4, 0.106450774, 0.786570725, 0.106450774, 2.63865e-4], rtol=1e-5)
    g = design.GriffinLimNodePlan(16, 10, 16000, 40, 8)
    assert (g.fft_size, g.hop, g.block_len, g.context_width, g.offsets) == (256, 160, 3, 1, [0, 160])
    assert len(g.lp_a) == 6 and abs(g.lp_a[1] - 4.87292) < 1e-4 and g.window[0] < 0


def test_scan_plan_modes(monkeypatch):
    fe = FeatureExtractor(2048)
    k, clen, hor, phi = fe.scan_plan(1228800, 4096)
    assert k == 2 and hor < clen and phi is None and hor % 1024 == 0            # balanced pieces with the modal tail
    assert fe.pieces_with_tail(1228800, 128) and not fe.pieces_with_tail(307200, 64)
    monkeypatch.setenv('SGS_FEAT_TAIL', '0')
    fe = FeatureExtractor(2048)
    k, clen, hor, phi = fe.scan_plan(1228800, 4096)
    assert k == 3 and hor < clen and phi is None and hor % 1024 == 0            # truncated zero-state pass
    k, clen, hor, phi = fe.scan_plan(307200, 64)
    assert k > 2 and hor == clen and phi is not None and phi.shape == (48, 48)  # exact carry through Phi
    assert fe.scan_plan(1400, 6)[0] == 1
    # Phi by sequential propagation reproduces the zero-input response of the cascade
    A = fe.transition()
    s = np.random.default_rng(0).normal(size=48)
    t = s.copy()
    for _ in range(clen):
        t = A @ t
    assert np.abs(phi @ s - t).max() <= 1e-9 * np.abs(t).max()


def test_node_runtime_contract():
    from livenodes import Node, LambdaNode, ChannelSelector
    src = Node.Node(name='src', has_inputs=False)
    sel = ChannelSelector.ChannelSelector(exclude=[1])(src)
    dbl = LambdaNode.LambdaNode(lambda f: f * 2)(sel)
    got = []
    dbl.add_output(got.append)
    src.output_data(np.arange(6.0).reshape(2, 3))
    assert np.array_equal(got[0], [[0, 4], [6, 10]])
    with pytest.raises(ValueError):
        dbl.set_inputs(src)                                  # input already set
    with pytest.raises(ValueError):
        Node.Node(has_inputs=False)(src)                     # module has no inputs
    sink = Node.Node(has_outputs=False)
    with pytest.raises(ValueError):
        sink.add_output(got.append)
    facade = Node.Node(name='facade')
    facade.set_passthrough(sel, dbl)
    assert facade.add_data == sel.add_data and facade.add_output == dbl.add_output


def test_framebuffer_framing_matches_oracle_schedule():
    from livenodes import FrameBuffer
    fb = FrameBuffer.FrameBuffer(21, 1, 1000, warm_start=True)
    rows = []
    fb.add_output(lambda f: rows.append(f.copy()))
    x = np.arange(1.0, 41.0).reshape(-1, 1)
    for v in x:
        fb.add_data(v.reshape(1, 1))
    assert len(rows) == 40 and rows[0].shape == (21, 1)
    assert rows[0][-1, 0] == 1.0 and np.all(rows[0][:-1] == 0)          # 20 warm-start zeros then the first sample
    assert np.array_equal(rows[25][:, 0], np.arange(6.0, 27.0))
    stacked = np.array([r[::5, 0] for r in rows])
    assert np.array_equal(stacked, O.stack_online(x)[:, :5])


def test_channel_selector_equals_np_delete():
    from livenodes import ChannelSelector
    x = np.arange(40.0).reshape(5, 8)
    for bad in ([], [0], [2, 5], [7, 1, 1], [-1]):
        got = []
        sel = ChannelSelector.ChannelSelector(exclude=bad)
        sel.add_output(got.append)
        sel.add_data(x)
        sel.add_data(x[:, :6] if not bad or max(bad) < 6 and min(bad) >= -6 else x)
        assert np.array_equal(got[0], np.delete(x, bad, axis=1))
        assert got[0] is not x and not np.shares_memory(got[0], x)      # a fresh array per chunk, as in the reference
        y = x[:, :6] if not bad or max(bad) < 6 and min(bad) >= -6 else x
        assert np.array_equal(got[1], np.delete(y, bad, axis=1))


def test_receiver_batches_and_flushes():
    from livenodes import Receiver
    rec = Receiver.Receiver(flush_interval=3600.0)
    for i in range(5):
        rec.add_data(np.full(3, i))
    assert len(rec.data) == 0                                 # nothing sent to the Manager yet
    got = rec.get_data()                                      # same-process read hands the batch over first
    assert [int(g[0]) for g in got] == [0, 1, 2, 3, 4] and len(rec.data) == 5
    rec.add_data(np.full(3, 5))
    rec.stop_processing()
    assert len(rec.data) == 6
    assert len(rec.get_data(clear=True)) == 6 and rec.get_data() == []
    eager = Receiver.Receiver(flush_interval=0, perform_timing=True)
    eager.add_data(1.5)
    assert len(eager.data) == 1 and eager.get_data()[0][1] == 1.5


def _feed_receiver(rec, n):
    for i in range(n):
        rec.add_data(i)


def test_receiver_flushes_when_a_forked_feeder_exits():
    """The reference's execution model: frames are appended in a forked child and read by the parent afterwards."""
    import multiprocessing
    from livenodes import Receiver
    rec = Receiver.Receiver(flush_interval=3600.0)
    rec.add_data(-1)                                           # the parent's own batch must not be replayed by the child
    ctx = multiprocessing.get_context('fork')
    p = ctx.Process(target=_feed_receiver, args=(rec, 7))
    p.start(); p.join()
    assert p.exitcode == 0
    assert sorted(rec.get_data()) == [-1, 0, 1, 2, 3, 4, 5, 6]


def test_fused_chain_is_found_only_for_the_reference_wiring():
    """sgs.chain.find_chain recognises decode.setup_decoder's graph (decode.py:152-183) and nothing else; no device needed."""
    import pickle
    import decode
    from livenodes import Node, LambdaNode, Dequantization, LDASynthesis, ECogFeatCalc, GriffinLim
    from sgs import chain
    from sgs.training import PackedLDA
    ests = []
    for _ in range(3):
        e = PackedLDA(); e.coef_ = np.zeros((9, 4)); e.intercept_ = np.zeros(9); e.classes_ = np.arange(9.0)
        ests.append(e)
    blob, med = pickle.dumps(ests), np.zeros((3, 9))
    src = Node.Node(name='src', has_inputs=False)
    decode.setup_decoder(src, 1024, blob, med, [], np.arange(4), include_soundcard=False, nb_mel_bins=3)
    feat = src.output_classes[0].output_classes[0]
    found = chain.find_chain(feat)
    assert found is not None and [type(n).__name__ for n in found] == ['LDASynthesis', 'Dequantization', 'GriffinLimSynthesis']
    os_env = __import__('os').environ
    os_env['SGS_FUSED_CHAIN'] = '0'
    try:
        assert chain.find_chain(feat) is None
    finally:
        del os_env['SGS_FUSED_CHAIN']
    # a node between LDA and dequantisation, or a second LDA consumer, is not the reference wiring
    src2 = Node.Node(name='src2', has_inputs=False)
    f2 = ECogFeatCalc.ECogFeatCalc(1024, 50, 10)(src2)
    l2 = LDASynthesis.LDASynthesis(blob, select=np.arange(4))(f2)
    mid = LambdaNode.LambdaNode(lambda f: f)(l2)
    Dequantization.Dequantization(med)(mid)
    assert chain.find_chain(f2) is None
    src3 = Node.Node(name='src3', has_inputs=False)
    f3 = ECogFeatCalc.ECogFeatCalc(1024, 50, 10)(src3)
    LDASynthesis.LDASynthesis(blob, select=np.arange(4))(f3)
    LDASynthesis.LDASynthesis(blob, select=np.arange(4))(f3)
    assert chain.find_chain(f3) is None


def test_artefacts_round_trip(tmp_path):
    """params / recording / decoding artefacts either side of a run (train.py:171-205, decode.py:186-219, 299-313): what
    train.store_training_to_file writes is what decode.load_params reads, and the decoding outputs come back bit-exact."""
    import configparser
    import pickle
    from scipy.io.wavfile import read as wavread
    import decode
    import train
    cfg = configparser.ConfigParser()
    cfg['General'] = {'storage_dir': str(tmp_path), 'session': 's1'}
    os.makedirs(tmp_path / 's1')
    rng = np.random.default_rng(0)
    medians, select, bad = rng.normal(size=(40, 9)), rng.permutation(640)[:150], np.array([3, 17])
    estimators = [{'coef_': rng.normal(size=(9, 150))} for _ in range(3)]
    train.store_training_to_file(cfg, rng.normal(size=(50, 150)), rng.normal(size=(50, 40)), medians, estimators, bad, select)
    blob, m2, b2, s2 = decode.load_params(str(tmp_path / 's1'))
    assert np.array_equal(m2, medians) and np.array_equal(b2, bad) and np.array_equal(s2, select)
    back = pickle.loads(blob)
    assert all(np.array_equal(a['coef_'], b['coef_']) for a, b in zip(back, estimators))
    assert (tmp_path / 's1' / 'LDAs.pkl').exists() and (tmp_path / 's1' / 'train.ini').exists()

    spec, audio = rng.normal(size=(30, 40)), rng.integers(-32768, 32767, 4800).astype(np.int16)
    seeg = rng.normal(size=(614, 8)).astype(np.float32)
    with pytest.raises(ValueError):
        decode.store_decoding_to_file(spec, audio, seeg, 2048)                # reference signature; needs decode.run_dir
    decode.run_dir, decode.config = str(tmp_path), cfg                       # ... which the reference's __main__ sets as globals
    try:
        decode.store_decoding_to_file(spec, audio, seeg, 2048)
    finally:
        decode.run_dir = decode.config = None
    assert (tmp_path / 'params.h5').exists() is False and (tmp_path / 's1' / 'params.h5').exists() and (tmp_path / 'sEEG.hdf').exists()
    sr, a2 = wavread(str(tmp_path / 'audio.wav'))
    assert sr == 16000 and a2.dtype == np.int16 and np.array_equal(a2, audio)
    assert np.array_equal(np.load(tmp_path / 'spectrogram.npy'), spec)
    e2, sf = decode.load_seeg(str(tmp_path / 'sEEG.hdf'))
    assert sf == 2048 and np.array_equal(e2, seeg)
    assert (tmp_path / 'decode.ini').exists()


def test_exp_angle_tables_match_generator():
    """csrc/exp_angle.cuh carries the polynomial and the 8 x 65 constants that tools/gen_exp_angle.py derives with mpmath; the
    fp64 instruction sequence emulated there (every FMA rounded once) stays within 1.5 ulp of exp(atan2(y, x))."""
    import re
    import sys
    import mpmath as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'tools'))
    import gen_exp_angle as g
    hdr = open(os.path.join(root, 'closed-loop-seeg-speech-synthesis_b200', 'csrc', 'exp_angle.cuh')).read()
    P = g.cheb_interp_monomial(lambda s: mp.exp(mp.atan(s)), g.DEG, -g.S_MAX, g.S_MAX)
    body = hdr[hdr.index('kEaPoly[kEaDeg + 1] = {'):]
    got = re.findall(r'-?0x1\.[0-9a-f]+p[+-]\d+', body[:body.index('};')])
    assert got == [float(v).hex() for v in P]
    tab = hdr[hdr.index('kEaTab[kEaTabLen] = {'):]
    got = re.findall(r'-?0x1\.[0-9a-f]+p[+-]\d+', tab[:tab.index('};')])
    want = [float(g.case_constant(c & 1, (c >> 1) & 1, (c >> 2) & 1, j)).hex() for c in range(8) for j in range(g.GRID + 1)]
    assert got == want
    rng = np.random.default_rng(5)
    worst = 0.0
    for i in range(300):
        x, y = (float(v) for v in rng.normal(size=2) * 10.0 ** rng.uniform(-3, 3, 2))
        val, s = g.emulate(y, x, P, mp.mpf(2) ** -20 * (1 if i % 2 else -1))
        worst = max(worst, float(abs(val / mp.exp(mp.atan2(mp.mpf(y), mp.mpf(x))) - 1)))
        assert abs(s) <= float(g.S_MAX)
    assert worst < 1.5 * 2.2e-16


def test_receiver_holds_frames_until_asked_by_default():
    """No Manager round trip while frames stream (every frame over 1 ms of the latency leg was one, tools/latency_tail.py):
    the default Receiver hands over in get_data / stop_processing / flush, or once max_held_bytes have piled up."""
    from livenodes import Receiver
    rec = Receiver.Receiver()
    for i in range(10):
        rec.add_data(np.full(4, i))
    assert rec._thread is None and len(rec.data) == 0
    assert [int(a[0]) for a in rec.get_data()] == list(range(10))
    small = Receiver.Receiver(max_held_bytes=100)
    for i in range(5):
        small.add_data(np.zeros(8))                      # 64 bytes each: the second frame crosses the limit
    assert len(small.data) == 4 and len(small.get_data()) == 5


def test_receiver_timed_flush_runs_off_the_graph_thread():
    """Timed hand-overs go through the flusher thread; order is kept and a synchronous flush drains the queue first."""
    import threading
    from livenodes import Receiver
    rec = Receiver.Receiver(flush_interval=0.01)
    for i in range(60):
        rec.add_data(i)
        time.sleep(0.002)
    assert rec._thread is not None and rec._thread.name.endswith('-flusher') and rec._thread is not threading.current_thread()
    assert rec.get_data() == list(range(60))
    rec.add_data(60)
    rec.stop_processing()
    assert list(rec.data) == list(range(61))


def test_gl_write_head_table_equals_the_node_loop():
    """GriffinLimNodeOp.positions (cumsum, cached) reproduces the node's per-frame float accumulation
    `outputBufferPosMs += frameShiftMs; pos = int(ms / 1000 * sampleRate)` (GriffinLim.py:115-120) bit for bit."""
    from sgs.griffinlim import GriffinLimNodeOp
    for shift in (10, 10.0, 9.97):
        op = GriffinLimNodeOp(16, shift, 16000, 40)
        for start in (0.0, 3.3):
            got = op.positions(70000, start)
            ms, want = start, np.empty(70000, dtype=np.int32)
            for k in range(70000):
                ms += op.plan.frame_shift_ms
                want[k] = int((ms / 1000.0) * op.plan.sample_rate)
            assert np.array_equal(got, want)
            assert op.positions(70000, start) is got and not got.flags.writeable       # cached, shared, read-only


def test_receiver_keeps_the_tail_when_the_feeder_is_terminated():
    """Sender.stop_processing ends its feeder with Process.terminate() (reference Sender.py:64-70): SIGTERM skips the exit
    finalizers, so the batching Receiver hands its frames over from a SIGTERM hook - like the reference's per-frame append
    (flush_interval=0, attached next to it) nothing the feeder produced is lost, and the order is kept."""
    from livenodes import Receiver, Sender
    data = np.arange(4000, dtype=np.float64)[:, None]
    snd = Sender.Sender(data, 1000, 2, asap=False)                       # 2-sample frames every 2 ms: runs for 4 s if not stopped
    batched = Receiver.Receiver(flush_interval=3600.0)(snd)             # would only ever flush at exit
    timed = Receiver.Receiver(flush_interval=0.25)(snd)
    eager = Receiver.Receiver(flush_interval=0)(snd)                    # the reference's behaviour
    snd.start_processing()
    time.sleep(0.7)
    snd.stop_processing()
    a, b, c = batched.get_data(), timed.get_data(), eager.get_data()
    assert len(c) > 100                                                 # the feeder did run
    # the signal lands between two receivers' callbacks at worst: lengths differ by at most the frame in flight
    assert abs(len(a) - len(c)) <= 1 and abs(len(b) - len(c)) <= 1, (len(a), len(b), len(c))
    for got in (a, b):
        first = np.array([f[0, 0] for f in got])
        assert np.array_equal(first, 2.0 * np.arange(len(got)))         # complete and in order


def test_receiver_resends_failed_batches_in_order():
    from livenodes import Receiver

    class Flaky(list):
        fail = 2

        def extend(self, batch):
            if Flaky.fail > 0:
                Flaky.fail -= 1
                raise ConnectionError('manager away')
            list.extend(self, batch)

    rec = Receiver.Receiver(flush_interval=0.01)
    rec.data = Flaky()
    for i in range(40):
        rec.add_data(i)
        time.sleep(0.002)
    rec.flush()
    assert list(rec.data) == list(range(40))


@pytest.mark.parametrize('sr,shift_ms', [(2048, 10), (512, 10), (2048, 0.5), (1024, 1), (600, 2.5)])
def test_feature_node_schedule_never_defers_a_frame(sr, shift_ms):
    """ECogFeatCalc._schedule (host side of the streaming push): every frame is scheduled in the push whose samples complete
    it, at most 16 per push - pushes are cut short when more would end inside 128 samples - and the ends follow
    FrameBuffer.py:35,177 exactly.  No device needed."""
    from livenodes import ECogFeatCalc
    node = ECogFeatCalc.ECogFeatCalc(sr, 50, shift_ms, has_inputs=False)
    plan = node._fe.plan
    first_ms = (float(plan.frame_size) / float(sr)) * 1000.0
    want = lambda k: round(((first_ms + k * float(shift_ms)) / 1000.0) * float(sr)) - plan.zero_fill
    total, k = 0, 0
    left = 5000
    while left > 0:
        n, ends, idx = node._schedule(min(128, left))
        assert 1 <= n <= 128 and len(ends) <= 16
        for e, i in zip(ends, idx):
            assert i == k and e == want(k)
            assert total < e <= total + n or (k == 0 and e <= total + n)          # ends inside this push's samples
            k += 1
        assert want(k) > total + n                                               # the next frame is not complete yet
        node._consumed += n
        total += n
        left -= n
    assert k > 5000 * 1000.0 / sr / shift_ms - 10


def _params_like(rng):
    import pickle
    return {'bad_channels': np.array([3, 17]), 'medians_array': rng.normal(size=(40, 9)),
            'estimators': np.void(pickle.dumps([{'coef_': rng.normal(size=(9, 150))} for _ in range(3)])),
            'select': rng.permutation(640)[:150], 'sEEG': rng.normal(size=(614, 8)).astype(np.float32), 'sEEG_sr': np.int32(2048),
            'no_bad_channels': np.array([]), 'u16': np.arange(7, dtype=np.uint16), 'i8_3d': rng.integers(-100, 100, (2, 3, 4)).astype(np.int8),
            'scalar_f64': np.float64(3.25)}


def test_hdf5lite_round_trip_and_file_structure(tmp_path):
    """sgs/hdf5lite.py (the h5py-free reader / writer of params.h5 and sEEG.hdf): every dataset comes back with its dtype,
    shape and bits, and the file has the structure the HDF5 specification prescribes for a version-0 superblock with a
    symbol-table root group - signature, 96-byte superblock, end-of-file address, TREE / HEAP / SNOD blocks at the addresses
    the superblock's root entry caches, entries in name order, 8-byte alignment of every object."""
    import struct
    from sgs import hdf5lite
    rng = np.random.default_rng(5)
    data = _params_like(rng)
    path = str(tmp_path / 'params.h5')
    hdf5lite.write(path, data)
    back = hdf5lite.read(path)
    assert sorted(back) == sorted(data)
    for k, v in data.items():
        got = back[k]
        if isinstance(v, np.void):
            assert isinstance(got, np.void) and got.tobytes() == v.tobytes()
        else:
            v = np.asarray(v)
            assert np.asarray(got).dtype == v.dtype and np.asarray(got).shape == v.shape and np.array_equal(got, v), k
    assert list(hdf5lite.read(path, ['select', 'sEEG_sr'])) == ['select', 'sEEG_sr']
    with pytest.raises(KeyError):
        hdf5lite.read(path, ['missing'])
    raw = open(path, 'rb').read()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n' and raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, internal_k, flags = struct.unpack_from('<HHI', raw, 16)
    base, free, eof, driver = struct.unpack_from('<QQQQ', raw, 24)
    assert (leaf_k, internal_k, flags, base) == (4, 16, 0, 0) and free == driver == 2 ** 64 - 1 and eof == len(raw)
    name_off, root_hdr, cache, _, btree, heap = struct.unpack_from('<QQIIQQ', raw, 56)
    assert (name_off, root_hdr, cache) == (0, 96, 1)
    assert raw[btree:btree + 4] == b'TREE' and raw[heap:heap + 4] == b'HEAP'
    # the root object header carries the same two addresses in its symbol-table message (type 0x0011)
    version, n_msgs, refs, size = struct.unpack_from('<BxHII', raw, root_hdr)
    mtype, msize = struct.unpack_from('<HH', raw, root_hdr + 16)
    assert (version, n_msgs, refs, mtype, msize) == (1, 1, 1, 0x11, 16) and struct.unpack_from('<QQ', raw, root_hdr + 24) == (btree, heap)
    ntype, level, used, left, right, key0 = struct.unpack_from('<BBHQQQ', raw, btree + 4)
    assert (ntype, level, used, key0) == (0, 0, 2, 0) and left == right == 2 ** 64 - 1          # 10 datasets = two leaf nodes of <= 8
    seg_size, free_head, seg = struct.unpack_from('<QQQ', raw, heap + 8)
    assert free_head == 1 and seg % 8 == 0 and raw[seg:seg + 8] == bytes(8)
    names = []
    for i in range(used):
        child, key = struct.unpack_from('<QQ', raw, btree + 32 + 16 * i)
        assert raw[child:child + 4] == b'SNOD' and child % 8 == 0
        n = struct.unpack_from('<H', raw, child + 6)[0]
        for j in range(n):
            off, hdr = struct.unpack_from('<QQ', raw, child + 8 + 40 * j)
            assert hdr % 8 == 0 and raw[hdr] == 1
            names.append(raw[seg + off:raw.index(b'\x00', seg + off)].decode())
        assert raw[seg + key:raw.index(b'\x00', seg + key)].decode() == names[-1]             # key = largest name of the child to its left
    assert names == sorted(data, key=lambda s: s.encode())
    with pytest.raises(hdf5lite.Hdf5LiteError):
        hdf5lite.write(str(tmp_path / 'bad.h5'), {'strings': np.array(['a', 'b'])})
    (tmp_path / 'junk.h5').write_bytes(b'not an hdf5 file' * 100)
    with pytest.raises(hdf5lite.Hdf5LiteError):
        hdf5lite.read(str(tmp_path / 'junk.h5'))


def test_hdf5lite_interchanges_with_h5py_when_present(tmp_path):
    """Runs only where h5py is importable (not in the build image): files written by either side are read by the other."""
    h5py = pytest.importorskip('h5py')
    from sgs import hdf5lite
    rng = np.random.default_rng(6)
    data = _params_like(rng)
    ours, theirs = str(tmp_path / 'ours.h5'), str(tmp_path / 'theirs.h5')
    hdf5lite.write(ours, data)
    with h5py.File(ours, 'r') as hf:
        for k, v in data.items():
            got = hf[k][()]
            assert (got.tobytes() == v.tobytes()) if isinstance(v, np.void) else np.array_equal(got, v), k
    with h5py.File(theirs, 'w') as hf:
        for k, v in data.items():
            hf.create_dataset(k, data=v)
    back = hdf5lite.read(theirs)
    for k, v in data.items():
        assert (back[k].tobytes() == v.tobytes()) if isinstance(v, np.void) else np.array_equal(back[k], v), k


def test_realtime_scheduling_is_restored():
    """decode.realtime pins the calling thread and asks for SCHED_FIFO; whatever it was granted, affinity and policy are back
    afterwards (and a refusal is reported, not raised)."""
    import decode
    before = (os.sched_getaffinity(0), os.sched_getscheduler(0))
    with decode.realtime() as rt:
        assert set(rt.applied) == {'pinned_to', 'policy'}
        if isinstance(rt.applied['pinned_to'], int):
            assert os.sched_getaffinity(0) == {rt.applied['pinned_to']}
    assert (os.sched_getaffinity(0), os.sched_getscheduler(0)) == before
