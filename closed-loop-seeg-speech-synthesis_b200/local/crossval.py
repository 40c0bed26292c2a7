"""Many-fold re-train / decode driver: the compute core of the reference's evaluation experiment 1
(eval_steps/exp1.py:26-38 train_decode_worker, :54-101 fold construction, :103-160 runs), without its session files,
plots and wav export.

Every fold trains a fresh model on the recording with the held-out stretch cut out (train.train) and decodes the held-out
stretch with it (decode.perform_offline_decoding); `randomize=True` breaks the alignment between neural and audio data
by rotating the training recording at a random sample, which is how the reference estimates chance level."""
import logging
import pickle

import numpy as np

logger = logging.getLogger('crossval.py')


def train_decode_worker(k, x_train, y_train, x_test, y_test, eeg_sr, audio_sr, bad_channels, norm_factor):
    """One fold (eval_steps/exp1.py:26-38): returns (k, reconstructed spectrogram, reference spectrogram, decoded audio)."""
    from train import train
    from decode import perform_offline_decoding
    logger.info('Processing Fold k={}'.format(k))
    _, _, medians, estimators, select = train(x_train, y_train, eeg_sr, audio_sr, bad_channels)
    params = (pickle.dumps(estimators), medians, bad_channels, select)
    reco_spec, out_audio, _, _ = perform_offline_decoding(params, x_test, eeg_sr, norm_factor)
    logger.info('Finished Fold k={}, shape: {}'.format(k, reco_spec.shape))
    return k, reco_spec, y_test, out_audio


def construct_folds(eeg, audio, eeg_sr, audio_sr, bad_channels, norm_factor, nb_folds=10, randomize=False, rng=None):
    """Argument tuples of train_decode_worker for `nb_folds` contiguous held-out stretches (exp1.py:54-101; the reference
    cuts at word boundaries of its session file, here the recording is cut into equal stretches of whole 10 ms frames)."""
    from local.offline import compute_spectrogram
    from sgs.spectrogram import decimate
    eeg, audio = np.asarray(eeg), np.asarray(audio, dtype=np.float64)
    seconds = min(len(eeg) / float(eeg_sr), len(audio) / float(audio_sr))
    n_frames = int(seconds * 100)
    q = int(round(audio_sr / 16000))
    rng = np.random.default_rng() if rng is None else rng
    arguments = []
    for k in range(1, nb_folds + 1):
        f0, f1 = (k - 1) * n_frames // nb_folds, k * n_frames // nb_folds
        e0, e1 = int(f0 / 100.0 * eeg_sr), int(f1 / 100.0 * eeg_sr)
        a0, a1 = int(f0 / 100.0 * audio_sr), int(f1 / 100.0 * audio_sr)
        x_train = np.vstack([eeg[:e0], eeg[e1:int(seconds * eeg_sr)]]).astype(np.float64)
        y_train = np.concatenate([audio[:a0], audio[a1:int(seconds * audio_sr)]])
        x_test = eeg[e0:e1]
        held_out = audio[a0:a1]
        y_test = compute_spectrogram(decimate(held_out, q) if q > 1 else held_out, window_length=0.016)
        if randomize:
            r = int(rng.integers(0, len(x_train)))
            logger.info('Random splitting at index {}'.format(r))
            x_train = np.vstack([x_train[r:], x_train[:r]])
        arguments.append((k, x_train, y_train, x_test, y_test, eeg_sr, audio_sr, bad_channels, norm_factor))
    return arguments


def cross_validate(eeg, audio, eeg_sr, audio_sr, bad_channels, norm_factor=10, nb_folds=10, randomize=False, rng=None):
    """Runs all folds; returns (reconstructed, reference) spectrograms of the whole recording stacked in time, the decoded
    audio, and the per-bin Pearson correlations (mean, std, list) between reconstruction and reference."""
    from local.offline import pearson_correlation
    results = [train_decode_worker(*args) for args in
               construct_folds(eeg, audio, eeg_sr, audio_sr, bad_channels, norm_factor, nb_folds, randomize, rng)]
    results.sort(key=lambda r: r[0])
    reco, orig, wav = [], [], []
    for _, r, o, w in results:
        n = min(len(r), len(o))                              # the streaming framing emits a few frames fewer than the target has
        reco.append(r[:n]); orig.append(np.asarray(o)[:n]); wav.append(w)
    reco, orig = np.vstack(reco), np.vstack(orig)
    return reco, orig, np.hstack(wav), pearson_correlation(orig, reco, return_means=True)
