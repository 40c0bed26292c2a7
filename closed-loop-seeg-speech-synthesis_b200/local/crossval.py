"""Many-fold re-train / decode driver: the compute core of the reference's evaluation experiment 1
(eval_steps/exp1.py:26-38 train_decode_worker, :54-101 fold construction, :103-160 runs), without its session files,
plots and wav export.

Every fold trains a fresh model on the recording with the held-out stretch cut out (train.train) and decodes the held-out
stretch with it (decode.perform_offline_decoding); `randomize=True` breaks the alignment between neural and audio data
by rotating the training recording at a random sample, which is how the reference estimates chance level.

The reference runs its 10 folds (x 100 randomised runs) one after the other on one core.  Here
  * the recording and the audio are uploaded ONCE; a fold's training set is a device-to-device concatenation of the two
    stretches either side of the held-out one, its model goes from train to decode without leaving the process, and the
    held-out stretch is decoded from the resident tensor - per fold only the fitted coefficients and the decoded results
    cross the bus;
  * folds are independent, so under torch.distributed (one process per GPU) they are dealt over the ranks with
    decode.session_shard - no collective on the data path; the results are all-gathered at the end.
The features of a fold are computed over the spliced recording exactly as the reference does (its filters run across the
splice, exp1.py:82-84), so per-fold results equal train_decode_worker's on the same arrays."""
import logging
import pickle
import time

import numpy as np

logger = logging.getLogger('crossval.py')

last_profile = {}       # wall seconds of the stages of the last cross_validate in this process


def train_decode_worker(k, x_train, y_train, x_test, y_test, eeg_sr, audio_sr, bad_channels, norm_factor):
    """One fold with the reference's signature and host arrays (eval_steps/exp1.py:26-38): returns (k, reconstructed
    spectrogram, reference spectrogram, decoded audio)."""
    from train import train
    from decode import perform_offline_decoding
    logger.info('Processing Fold k={}'.format(k))
    _, _, medians, estimators, select = train(x_train, y_train, eeg_sr, audio_sr, bad_channels)
    params = (pickle.dumps(estimators), medians, bad_channels, select)
    reco_spec, out_audio, _, _ = perform_offline_decoding(params, x_test, eeg_sr, norm_factor)
    logger.info('Finished Fold k={}, shape: {}'.format(k, reco_spec.shape))
    return k, reco_spec, y_test, out_audio


def fold_bounds(n_eeg, n_audio, eeg_sr, audio_sr, nb_folds):
    """[(k, (e0, e1), (a0, a1), e_end, a_end)]: sample ranges of the `nb_folds` contiguous held-out stretches (exp1.py:54-101;
    the reference cuts at word boundaries of its session file, here the recording is cut into equal stretches of whole
    10 ms frames)."""
    seconds = min(n_eeg / float(eeg_sr), n_audio / float(audio_sr))
    n_frames = int(seconds * 100)
    out = []
    for k in range(1, nb_folds + 1):
        f0, f1 = (k - 1) * n_frames // nb_folds, k * n_frames // nb_folds
        out.append((k, (int(f0 / 100.0 * eeg_sr), int(f1 / 100.0 * eeg_sr)), (int(f0 / 100.0 * audio_sr), int(f1 / 100.0 * audio_sr)),
                    int(seconds * eeg_sr), int(seconds * audio_sr)))
    return out


def construct_folds(eeg, audio, eeg_sr, audio_sr, bad_channels, norm_factor, nb_folds=10, randomize=False, rng=None):
    """Argument tuples of train_decode_worker (host arrays), as the reference builds them."""
    from local.offline import compute_spectrogram
    from sgs.spectrogram import decimate
    eeg, audio = np.asarray(eeg), np.asarray(audio, dtype=np.float64)
    q = int(round(audio_sr / 16000))
    rng = np.random.default_rng() if rng is None else rng
    arguments = []
    for k, (e0, e1), (a0, a1), e_end, a_end in fold_bounds(len(eeg), len(audio), eeg_sr, audio_sr, nb_folds):
        x_train = np.vstack([eeg[:e0], eeg[e1:e_end]]).astype(np.float64)
        y_train = np.concatenate([audio[:a0], audio[a1:a_end]])
        x_test = eeg[e0:e1]
        held_out = audio[a0:a1]
        y_test = compute_spectrogram(decimate(held_out, q) if q > 1 else held_out, window_length=0.016)
        if randomize:
            r = int(rng.integers(0, len(x_train)))
            logger.info('Random splitting at index {}'.format(r))
            x_train = np.vstack([x_train[r:], x_train[:r]])
        arguments.append((k, x_train, y_train, x_test, y_test, eeg_sr, audio_sr, bad_channels, norm_factor))
    return arguments


def _fold_on_device(k, eeg_d, audio_d, bounds, eeg_sr, audio_sr, norm_factor, rotate_at, seed):
    """One fold from the resident recording: (k, reconstructed spectrogram, reference spectrogram, audio) as host arrays."""
    import torch
    import decode
    from local.offline import compute_spectrogram
    from sgs import training
    from sgs.spectrogram import decimate
    _, (e0, e1), (a0, a1), e_end, a_end = bounds
    x_train = torch.cat([eeg_d[:e0], eeg_d[e1:e_end]])
    y_train = torch.cat([audio_d[:a0], audio_d[a1:a_end]])
    if rotate_at is not None:
        x_train = torch.cat([x_train[rotate_at:], x_train[:rotate_at]])
    _, _, medians, estimators, select = training.sharded_fit(x_train, y_train, eeg_sr, audio_sr, distributed=False)
    del x_train, y_train
    q = int(round(audio_sr / 16000))
    held_out = audio_d[a0:a1]
    y_test = compute_spectrogram(decimate(held_out, q) if q > 1 else held_out, window_length=0.016)
    dec = decode.OfflineDecoder(estimators, medians, select, eeg_sr, norm_factor, packet_size=32)
    spec, audio = dec.decode(eeg_d[e0:e1].contiguous(), None, seed)
    return k, spec.cpu().numpy(), y_test.cpu().numpy(), audio.cpu().numpy()


def cross_validate(eeg, audio, eeg_sr, audio_sr, bad_channels, norm_factor=10, nb_folds=10, randomize=False, rng=None, seed=0):
    """Runs all folds; returns (reconstructed, reference) spectrograms of the whole recording stacked in time, the decoded
    audio, and the per-bin Pearson correlations (mean, std, list) between reconstruction and reference.  Every rank of a
    torch.distributed job returns the same result.  The Griffin-Lim start noise comes from the device generator (seed, fold)."""
    import torch
    import decode
    from local.offline import pearson_correlation
    from sgs import _lib, hostio, training
    _lib.ensure_init()
    t_all = time.perf_counter()
    eeg = np.asarray(eeg)
    if eeg.dtype not in (np.float32, np.float64):
        eeg = eeg.astype(np.float64)
    if len(bad_channels) > 0:
        mask = np.ones(eeg.shape[1], bool)
        mask[np.asarray(bad_channels, dtype=int)] = False
        eeg = eeg[:, mask]
    t0 = time.perf_counter()
    eeg_d = hostio.upload(eeg)
    audio_d = hostio.upload(np.asarray(audio), np.float64)
    torch.cuda.synchronize()
    prof = {'upload_s': time.perf_counter() - t0}
    bounds = fold_bounds(len(eeg), len(audio), eeg_sr, audio_sr, nb_folds)
    rng = np.random.default_rng() if rng is None else rng
    rotate = [int(rng.integers(0, b[3] - (b[1][1] - b[1][0]))) if randomize else None for b in bounds]     # drawn on every rank alike
    dist, rank, world = training._dist(None)
    lo, hi = decode.session_shard(nb_folds, rank, world)
    t0 = time.perf_counter()
    mine = [_fold_on_device(b[0], eeg_d, audio_d, b, eeg_sr, audio_sr, norm_factor, rotate[i], seed + b[0])
            for i, b in enumerate(bounds) if lo <= i < hi]
    torch.cuda.synchronize()
    prof['folds_s'] = time.perf_counter() - t0
    prof['folds_on_this_rank'] = hi - lo
    t0 = time.perf_counter()
    results = [r for part in training._allgather_objects(mine) for r in part]
    prof['gather_s'] = time.perf_counter() - t0
    results.sort(key=lambda r: r[0])
    reco, orig, wav = [], [], []
    for _, r, o, w in results:
        n = min(len(r), len(o))                              # the streaming framing emits a few frames fewer than the target has
        reco.append(r[:n]); orig.append(np.asarray(o)[:n]); wav.append(w)
    reco, orig = np.vstack(reco), np.vstack(orig)
    prof['total_s'] = time.perf_counter() - t_all
    last_profile.clear()
    last_profile.update(prof)
    return reco, orig, np.hstack(wav), pearson_correlation(orig, reco, return_means=True)
