"""Decoding entry points with the reference's signatures (decode.py:71-183).

    setup_decoder               wires the livenodes graph exactly as the reference does (decode.py:152-183);
                                every signal-processing node is backed by a device stream.
    perform_offline_decoding    the reference pushes the whole recording through that graph from a forked
                                Sender (decode.py:71-96); here the same result is produced by the batched
                                kernels in one pass (features with the node's framing, LDA, dequantise +
                                smoothing, node-semantics Griffin-Lim), which is what makes many-session
                                throughput possible.  Set streaming=True to run the node graph instead.
    decode_sessions             many equally long sessions at once, device-resident (the throughput path).
"""
import logging
import pickle

import numpy as np

from livenodes import LDASynthesis, ECogFeatCalc, GriffinLim, Receiver, ChannelSelector, Sender, Dequantization
from sgs.features import FeatureExtractor
from sgs.griffinlim import GriffinLimNodeOp
from sgs.lda import LdaDecoder

logger = logging.getLogger('decode.py')


class OfflineDecoder:
    """Model + configuration bound to device plans; reusable across recordings of the same shape."""

    def __init__(self, estimators, medians_array, select, sfreq, gl_norm=10, packet_size=32, nb_mel_bins=40,
                 gl_iterations=8):
        if isinstance(estimators, (bytes, bytearray, np.void)):
            estimators = pickle.loads(bytes(estimators) if not isinstance(estimators, np.void) else estimators.tobytes())
        self.sfreq = sfreq
        self.packet_size = packet_size
        self.features = FeatureExtractor(sfreq, frame_len_ms=50, frame_shift_ms=10, model_order=4, step_size=5)
        self.lda = LdaDecoder(estimators, np.asarray(select), np.asarray(medians_array))
        self.gl = GriffinLimNodeOp(16, 10, 16000, nb_mel_bins, gl_iterations, 7900, gl_norm)

    def decode(self, eeg, noise=None, seed=0):
        """eeg: (T, C) or (S, T, C) with bad channels already removed; numpy or torch-CUDA.
        Returns (spectrogram (.., F, 40) float64, audio (.., n) int16)."""
        lp = self.features.log_power(eeg, online=True, chunk_size=self.packet_size)
        _, spec = self.lda.decode(lp, order=4, step=5, first_row=0, smooth=True, want_labels=False)
        audio = self.gl.synthesize(spec, noise, seed)
        return spec, audio


def decode_sessions(decoder, eeg_sessions, seed=0):
    """Throughput path: (S, T, C) device tensor in, (S, F, 40) spectrogram + (S, n) int16 audio out, all resident."""
    return decoder.decode(eeg_sessions, None, seed)


def _reference_noise(n_frames, block=480, first=1):
    """The np.random.rand(480) draws the node chain would make (GriffinLim.py:90): one per frame from the second
    frame on, taken from a COPY of numpy's global state - the reference draws them in a forked child, so the
    parent's stream does not advance."""
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    noise = np.zeros((n_frames, block))
    for k in range(first, n_frames):
        noise[k] = rs.rand(block)
    return noise


def perform_offline_decoding(params, eeg, sfreq, gl_norm, streaming=False):
    estimators_serialized, medians_array, bad_channels, select = params[0], params[1], params[2], params[3]
    logger.info('Using a sampling rate of {} for the sEEG data.'.format(sfreq))
    eeg = np.asarray(eeg)
    if streaming:
        eeg_sender = Sender.Sender(eeg, sfreq, 16, asap=True, name='sEEG-File-Sender')
        rec_seeg, rec_spec, rec_audio = setup_decoder(eeg_sender, sfreq, estimators_serialized, medians_array,
                                                      bad_channels, select, gl_norm, include_soundcard=False)
        eeg_sender.start_processing()
        eeg_sender.wait_for_completion()
        spectrogram = np.array(rec_spec.get_data())
        output_audio = np.hstack(rec_audio.get_data())
        received_sEEG = np.vstack(np.array(rec_seeg.get_data()))
        logger.info('Decoding completed.')
        return spectrogram, output_audio, received_sEEG, sfreq

    good = np.delete(eeg, np.asarray(bad_channels, dtype=int), axis=1) if len(bad_channels) else eeg
    decoder = OfflineDecoder(estimators_serialized, medians_array, select, sfreq, gl_norm, packet_size=32)
    lp = decoder.features.log_power(good, online=True, chunk_size=32)
    noise = _reference_noise(lp.shape[0])
    _, spectrogram = decoder.lda.decode(lp, order=4, step=5, first_row=0, smooth=True, want_labels=False)
    output_audio = decoder.gl.synthesize(spectrogram, noise)
    logger.info('Decoding completed.')
    return spectrogram, output_audio, eeg.copy(), sfreq


def setup_decoder(eeg_sender, sfreq, estimators_serialized, medians_array, bad_channels, select, gl_norm=10,
                  packet_size=32, include_soundcard=True, nb_mel_bins=40):
    eeg_select = ChannelSelector.ChannelSelector(exclude=bad_channels, name='BadChannelsExclusion')(eeg_sender)
    eeg_node = ECogFeatCalc.ECogFeatCalc(sfreq, frame_len_ms=50, frame_shift_ms=10,
                                         model_order=4, step_size=5, chunk_size=packet_size)(eeg_select)
    lda_node = LDASynthesis.LDASynthesis(estimators_serialized, select=select)(eeg_node)
    deq_node = Dequantization.Dequantization(medians_array)(lda_node)
    logger.info('Amplifier packet size: {}'.format(packet_size))
    gl_node = GriffinLim.GriffinLimSynthesis(
        originalFrameSizeMs=16, frameShiftMs=10, sampleRate=16000, melCoeffCount=nb_mel_bins,
        numReconstructionIterations=8, normFactor=gl_norm)(deq_node)
    rec_seeg = Receiver.Receiver(name='EEG')(eeg_sender)
    rec_spec = Receiver.Receiver(name='Spectrogram')(deq_node)
    rec_audio = Receiver.Receiver(name='Audio')(gl_node)
    if include_soundcard:
        # loudspeaker sinks (JACK / PortAudio) are sound-hardware I/O and out of scope
        logger.info('No soundcard sink attached: audio is available from the returned Audio receiver.')
    return rec_seeg, rec_spec, rec_audio
