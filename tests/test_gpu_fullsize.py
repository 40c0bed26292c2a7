"""The BENCHED geometry against the oracle: 12 sessions x 128 ch x 600 s @ 2048 Hz, decoded with the reference-trained
128-channel model (tests/golden/model128.npz).

With >= 12 sessions of this shape sgs_feat_extract takes its default decomposition - the time lines of the 48 stream groups
laid end to end and cut into 4 x SM-count equal pieces (k_iir_pieces<.., PIPES = 4>, ring 128, window 102) - and
sgs_lda_decode the tcgen05 filter (k_lda_pack + k_lda_tc) with fp64 re-scoring of near-ties: exactly the kernels
bench.py times.  One WHOLE session in the middle of the batch (its four stream groups are cut by ~50 piece boundaries, and
pieces straddle the group boundaries either side) is compared with the CPU oracle: features 1e-9, class indices and
spectrogram bit-exact, int16 audio +-1 LSB; all twelve are compared with the (group x chunk) decomposition."""
import numpy as np
import pytest

import oracle as O
from sgs import synth, _lib
from helpers import model128

pytestmark = pytest.mark.gpu

SR, N_CH, SECONDS, N_SESS, CHECK = 2048, 128, 600.0, 12, 5


@pytest.fixture(scope='module')
def batch():
    import torch
    import decode
    model, select, medians, _ = model128()
    dec = decode.OfflineDecoder(model, medians, select, SR, gl_norm=10, packet_size=64)
    x_host = synth.seeg_session(300 + CHECK, N_CH, SR, SECONDS)                       # float32 (T x C), the session the oracle decodes
    xd = synth.seeg_sessions_device([300 + s for s in range(N_SESS)], N_CH, SR, SECONDS)
    xd[CHECK].copy_(torch.from_numpy(x_host))
    return dec, xd, x_host, (model, select, medians)


def test_default_pieces_path_and_tensor_core_lda_whole_session_vs_oracle(batch, monkeypatch):
    import torch
    dec, xd, x_host, ((W, b, cls), select, medians) = batch
    monkeypatch.delenv('SGS_FEAT_PIECES', raising=False)
    monkeypatch.delenv('SGS_FEAT_PIECES_P', raising=False)
    monkeypatch.delenv('SGS_LDA_TC', raising=False)

    _lib.profile_enable(True)
    lp = dec.features.log_power(xd, online=True, chunk_size=64)
    labels, spec = dec.lda.decode(lp, order=4, step=5, first_row=0, smooth=True)
    torch.cuda.synchronize()
    ran = {k: _lib.profile_read(k)[1] for k in ('iir_pieces_state', 'iir_pieces_feat', 'iir_state', 'iir_feat', 'lda_pack', 'lda_tc', 'lda')}
    _lib.profile_enable(False)
    # the decomposition bench.py times, not the small-job fallback
    assert ran['iir_pieces_feat'] == 1 and ran['iir_pieces_state'] == 1 and ran['iir_feat'] == 0 and ran['iir_state'] == 0, ran
    assert ran['lda_tc'] == 1 and ran['lda_pack'] == 1, ran
    rescored, total = dec.lda.last_rescored(), labels.shape[0] * labels.shape[1] * labels.shape[2]
    print('trained model: %d of %d (frame, bin) pairs re-scored in fp64 (%.3f %%)' % (rescored, total, 100.0 * rescored / total))
    assert lp.shape == (N_SESS, 60000, N_CH) and labels.shape == (N_SESS, 60000, 40)
    assert 0 < rescored < 0.02 * total                                                    # the tensor-core pass still filters

    # ---- the whole session CHECK against the CPU oracle ---------------------------------------------------------------
    feats = O.ecog_feat_calc(x_host.astype(np.float64), SR, 50, 10, 4, 5, 50, 64)    # (60000, 640) stacked rows
    got = dec.features.stack(lp[CHECK], online=True).cpu().numpy()
    assert got.shape == feats.shape
    err = np.abs(got - feats).max()
    print('features, whole session: max |diff| %.3g' % err)
    assert err < 1e-9
    want_labels, _ = O.lda_predict_packed(feats, W, b, cls, select)
    want_spec = O.dequantization_node(want_labels, medians)
    got_labels = labels[CHECK].cpu().numpy()
    assert np.array_equal(got_labels, want_labels), int((got_labels != want_labels).sum())
    assert np.array_equal(spec[CHECK].cpu().numpy(), want_spec)
    # the other sessions went through the same kernels: fp64 scoring of every frame gives the same labels everywhere
    monkeypatch.setenv('SGS_LDA_TC', '0')
    lab64, _ = dec.lda.decode(lp, order=4, step=5, first_row=0, smooth=True)
    monkeypatch.delenv('SGS_LDA_TC')
    assert torch.equal(lab64, labels)

    # ---- pieces == (group x chunk) decomposition on all twelve sessions ------------------------------------------------------
    monkeypatch.setenv('SGS_FEAT_PIECES', '0')
    _lib.profile_enable(True)
    lp_grid = dec.features.log_power(xd, online=True, chunk_size=64)
    torch.cuda.synchronize()
    assert _lib.profile_read('iir_feat')[1] == 1 and _lib.profile_read('iir_pieces_feat')[1] == 0
    _lib.profile_enable(False)
    monkeypatch.delenv('SGS_FEAT_PIECES')
    d = (lp_grid - lp).abs().max().item()
    print('pieces vs (group x chunk) grid, 12 sessions: max |diff| %.3g' % d)
    assert d < 1e-10                                                                # prefix-sum re-basing points differ

    # ---- audio: the first 150 s of the session through the node-semantics Griffin-Lim, and single blocks further on ------------
    n_head = 15000
    noise = np.random.RandomState(4200).rand(60000, 480)
    pcm, _, blk = dec.gl.synthesize(want_spec, noise, want_filtered=True, want_blocks=True)
    ref = O.GriffinLimNode(16, 10, 16000, 40, 8, norm_factor=10)
    want_pcm, _ = ref.synthesize(want_spec[:n_head], noise[:n_head])
    dpcm = np.abs(pcm[:len(want_pcm)].astype(int) - want_pcm.astype(int))
    print('int16 audio, first %d frames: max |diff| %d LSB, %.4f %% of samples differ' % (n_head, dpcm.max(), 100 * (dpcm > 0).mean()))
    assert dpcm.max() <= 1 and (dpcm > 0).mean() < 2e-3
    for k in (20000, 33333, 47111, 59999):
        wb = ref.block(want_spec[k - 1:k + 1], noise[k])
        assert np.abs(blk[k] - wb).max() <= 1e-11 * np.abs(wb).max()


def test_batch_decode_equals_session_alone(batch):
    """decode_sessions on the batch (device noise generator keyed by (seed, session)) == the same session decoded alone with
    its index: the pieces decomposition cuts sessions at different places in the two calls."""
    import torch
    import decode
    dec, xd, _, _ = batch
    spec12, audio12 = decode.decode_sessions(dec, xd, seed=11)
    spec1, _ = dec.decode(xd[CHECK:CHECK + 1], None, seed=11)
    assert torch.equal(spec1[0], spec12[CHECK])
    spec0, audio0 = dec.decode(xd[0:1], None, seed=11)
    assert torch.equal(spec0[0], spec12[0]) and torch.equal(audio0[0], audio12[0])
