"""Training entry points with the reference's signatures (train.py:78-205).

    train(eeg, audio, sfreq_eeg, sfreq_audio, bad_channels, nb_mel_bins=40)
        -> (x_train[:, select], q_spectrogram, medians, estimators, select)

Feature extraction, the log-mel target, quantisation, Spearman ranking and the LDA statistics run on the device;
the 40 small eigen-problems are solved on the host and wrapped into scikit-learn LinearDiscriminantAnalysis
objects (when scikit-learn is importable) so the pickled model is interchangeable with the reference's.
The 48 kHz -> 16 kHz decimation of the audio (train.py:125, scipy.signal.decimate) runs on the device too
(sgs.spectrogram.decimate)."""
import logging
import os
import pickle

import numpy as np

from local.offline import compute_spectrogram, herff2016_b
from local.quantization import dequantize_spectrogram  # noqa: F401  (re-exported like the reference)
from local.utils import benchmark
from sgs import training

logger = logging.getLogger('train.py')


@benchmark
def quantization(y_train, nb_intervals=8):
    """Quantize the logMel spectrogram."""
    medians, borders, q_spectrogram = training.quantization(y_train, nb_intervals)
    if isinstance(q_spectrogram, np.ndarray):
        present = np.stack([(q_spectrogram == k).any(axis=0) for k in range(nb_intervals)])
    else:                                                     # device-resident labels: nine small reductions, no download
        import torch
        present = torch.stack([(q_spectrogram == k).any(dim=0) for k in range(nb_intervals)]).cpu().numpy()
    for i in range(present.shape[1]):
        diff = np.nonzero(~present[:, i])[0]
        if diff.size > 0:
            logger.info('Spec_bin "{}" misses samples for interval index/indices "{}"'.format(i, str(diff)))
    return medians, borders, q_spectrogram


@benchmark
def feature_selection(x_train, y_train, nb_feats=150):
    """Feature selection using rank correlation with the frame-mean of the target."""
    return training.feature_selection(x_train, y_train, nb_feats)


@benchmark
def train_estimators(estimators, x_train, y_train):
    """Fits one LDA per mel bin.  Every bin shares x_train, so one pass gathers the sufficient statistics of all
    40 fits; `estimators` (the caller's un-fitted list) is filled in place like the reference does."""
    stats = training.distributed_lda_stats(x_train, np.arange(x_train.shape[1]), y_train)
    fitted = training.fit_from_stats(stats)
    for mel_bin in range(len(estimators)):
        estimators[mel_bin] = fitted[mel_bin]
        if (mel_bin + 1) % 5 == 0:
            logger.info('{:02d} LDAs fitted so far.'.format(mel_bin + 1))


@benchmark
def compute_features(eeg, sfreq_eeg, audio, audio_sr):
    x_train = herff2016_b(eeg, sfreq_eeg, 0.05, 0.01)
    if audio_sr != 16000:
        from sgs.spectrogram import decimate
        audio = decimate(audio, int(round(audio_sr / 16000)))
    y_train = compute_spectrogram(audio, 16000, 0.016, 0.01)
    return x_train, y_train


def train(eeg, audio, sfreq_eeg, sfreq_audio, bad_channels, nb_mel_bins=40):
    """Host arrays in, host arrays + fitted estimators out, as in the reference (train.py:132-168) - every stage in between
    (features, decimation, log-mel target, quantisation, Spearman ranking, column selection, LDA statistics) works on
    device-resident tensors.  With torch.distributed initialised (one process per GPU, every rank passing the same arrays)
    the job is sharded as sgs/training.py:sharded_fit describes - features + Spearman by channel block, LDA statistics by
    row with one all-reduce - and every rank returns the same model."""
    from sgs import _lib
    _lib.ensure_init()
    eeg = np.asarray(eeg)
    if eeg.dtype not in (np.float32, np.float64):
        eeg = eeg.astype(np.float64)
    if len(bad_channels) > 0:
        logger.info('EEG original shape: {} x {}'.format(*eeg.shape))
        mask = np.ones(eeg.shape[1], bool)
        mask[bad_channels] = False
        eeg = eeg[:, mask]
        logger.info('EEG truncated shape: {} x {}'.format(*eeg.shape))
    else:
        logger.info('No bad channels specified.')
    x_train, y_train, medians, estimators, select = training.sharded_fit(eeg, audio, sfreq_eeg, sfreq_audio, nb_mel_bins)
    logger.info('x_train: ' + str(tuple(x_train.shape)))
    logger.info('y_train: ' + str(tuple(y_train.shape)))
    for stage in ('features_s', 'target_quantization_s', 'spearman_s', 'lda_stats_s', 'eigen_solves_s'):
        logger.info('Finished stage [{}] in {:.4f} seconds.'.format(stage[:-2], training.last_profile.get(stage, 0.0)))
    from sgs import hostio
    return hostio.download(x_train), hostio.download(y_train), medians, estimators, select


def store_training_to_file(config, x_train, y_train, medians, estimators, bad_channels, select):
    """Writes LDAs.pkl, training_features.npy, params.h5 and train.ini (train.py:171-205; the plot is out of scope).
    params.h5 is written with h5py when it is importable, else by sgs/hdf5lite.py - the same four root-group datasets."""
    base = os.path.join(config['General']['storage_dir'], config['General']['session'])
    with open(os.path.join(base, 'LDAs.pkl'), 'wb') as fh:
        pickle.dump(estimators, fh)
    np.save(os.path.join(base, 'training_features.npy'), x_train)
    from decode import _write_datasets
    _write_datasets(os.path.join(base, 'params.h5'), {'bad_channels': np.asarray(bad_channels), 'medians_array': medians,
                                                      'estimators': np.void(pickle.dumps(estimators)), 'select': select})
    with open(os.path.join(base, 'train.ini'), 'w') as configfile:
        config.write(configfile)
    logger.info('Training completed.')
