"""Batch Griffin-Lim (offline.griffin_lim) and the log-mel target (compute_spectrogram) on the device."""
import numpy as np
import pytest

import oracle as O
from sgs import synth
from helpers import load, batch_noise

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('T', [12, 40])
def test_griffin_lim_against_reference_fixture(T):
    from local.offline import griffin_lim
    GL = load('griffinlim.npz')
    lm = GL['batch_logmel_T%d' % T]
    want = GL['batch_pcm_T%d' % T]
    np.random.seed(500 + T)                                   # default path draws from numpy's global stream like the reference
    got = griffin_lim(lm)
    assert got.dtype == np.int16 and got.shape == want.shape
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3


def test_griffin_lim_batch_float_and_iterations():
    from sgs.griffinlim import griffin_lim_batch
    med = synth.default_medians()
    T = 30
    lm = synth.logmel_utterances(3, T, med, seed=3300)
    noise = np.stack([batch_noise(700 + u, T) for u in range(3)])
    for iters in (1, 8, 32):                                  # BASELINE config 4 uses 32 iterations
        pcm, wave = griffin_lim_batch(lm, noise, num_iterations=iters, want_waveform=True)
        for u in range(3):
            wp, wf = O.griffin_lim_offline(lm[u], noise[u], num_iterations=iters, return_float=True)
            assert np.abs(wave[u] - wf).max() <= 1e-9 * np.abs(wf).max()
            assert np.abs(pcm[u].astype(int) - wp.astype(int)).max() <= 1


def test_compute_spectrogram_matches_oracle():
    from local.offline import compute_spectrogram
    audio = synth.audio_session(1, 6.0)
    got = compute_spectrogram(audio, 16000, 0.016, 0.01)
    want = O.compute_spectrogram(audio, 16000, 0.016, 0.01)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-9
    G = load('train_decode.npz')                              # reference-computed rows of the 24 s training session
    full = compute_spectrogram(synth.audio_session(1, float(G['dur'])), 16000, 0.016, 0.01)[20:-4]
    assert np.abs(full[:64] - G['y_spec_head']).max() < 1e-9


def test_config4_full_size_batch():
    """BASELINE config 4 at full size - 4096 utterances x 2 s (200 frames), 32 iterations - resident on the device: three
    utterances against the CPU oracle from the same start noise, utterance independence (the same utterances decoded
    alone give the same bits) and run-to-run determinism."""
    import torch
    from sgs.griffinlim import griffin_lim_batch
    U, T, iters = 4096, 200, 32
    med = torch.from_numpy(synth.default_medians(40, 9)).cuda()
    g = torch.Generator(device='cuda'); g.manual_seed(3000)
    idx = torch.randint(0, 9, (U, T, 40), device='cuda', generator=g)
    spec = torch.gather(med[None, None].expand(U, T, 40, 9), 3, idx[..., None])[..., 0].contiguous()
    noise = torch.rand((U, 160 * (T - 1) + 800), dtype=torch.float64, device='cuda', generator=g)
    pcm = griffin_lim_batch(spec, noise, num_iterations=iters)
    assert pcm.shape == (U, 160 * T) and pcm.dtype == torch.int16
    assert torch.equal(pcm, griffin_lim_batch(spec, noise, num_iterations=iters))
    pick = [0, 2047, 4095]
    alone = griffin_lim_batch(spec[pick].contiguous(), noise[pick].contiguous(), num_iterations=iters)
    assert torch.equal(alone, pcm[pick])
    assert int(pcm.abs().max(dim=1).values.min()) == 32767                  # every utterance is scaled to its own peak
    for u in pick[:2]:
        want = O.griffin_lim_offline(spec[u].cpu().numpy(), noise[u].cpu().numpy(), num_iterations=iters)
        d = np.abs(pcm[u].cpu().numpy().astype(int) - want.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 2e-3
