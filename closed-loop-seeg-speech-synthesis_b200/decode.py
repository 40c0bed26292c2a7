"""Decoding entry points with the reference's signatures (decode.py:71-183).

    setup_decoder               wires the livenodes graph exactly as the reference does (decode.py:152-183);
                                every signal-processing node is backed by a device stream.
    perform_offline_decoding    the reference pushes the whole recording through that graph from a forked
                                Sender (decode.py:71-96); here the same result is produced by the batched
                                kernels in one pass (features with the node's framing, LDA, dequantise +
                                smoothing, node-semantics Griffin-Lim), which is what makes many-session
                                throughput possible.  Set streaming=True to run the node graph instead.
    decode_sessions             many equally long sessions at once, device-resident (the throughput path).
    session_shard               which sessions this rank decodes when one process per GPU shares a job (no collective).
    load_params / load_seeg / store_decoding_to_file
                                the artefacts either side of a decoding run (decode.py:186-219, 299-313): params.h5, the
                                sEEG recording, audio.wav, spectrogram.npy, sEEG.hdf.  HDF5 goes through h5py when it is
                                importable and through the built-in flat-file reader / writer sgs/hdf5lite.py when not.
"""
import os
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

    def decode(self, eeg, noise=None, seed=0, sessions_per_batch=1, pinned_outputs=False):
        """eeg: (T, C) or (S, T, C) with bad channels already removed; numpy or torch-CUDA.
        Returns (spectrogram (.., F, 40) float64, audio (.., n) int16).

        Host input with several sessions is streamed: while one batch of sessions is being decoded, the next one is
        copied to the device and the previous results are copied back (two CUDA streams, double-buffered staging).
        pinned_outputs=True returns views of page-locked result buffers owned by this decoder (valid until the next
        call) instead of fresh arrays."""
        from sgs import _lib
        if _lib._is_torch(eeg) or eeg.ndim == 2 or noise is not None or eeg.shape[0] <= sessions_per_batch:
            lp = self.features.log_power(eeg, online=True, chunk_size=self.packet_size)
            _, spec = self.lda.decode(lp, order=4, step=5, first_row=0, smooth=True, want_labels=False)
            audio = self.gl.synthesize(spec, noise, seed)
            return spec, audio
        return self._decode_streamed(eeg, seed, sessions_per_batch, pinned_outputs)

    def _decode_streamed(self, eeg, seed, per_batch, pinned_outputs):
        import torch
        from sgs import _lib
        _lib.ensure_init()
        S, T, C = eeg.shape
        if eeg.dtype not in (np.float32, np.float64):
            eeg = eeg.astype(np.float32)
        host = torch.from_numpy(np.ascontiguousarray(eeg))
        dev = torch.device('cuda', torch.cuda.current_device())
        copy_in, copy_out, compute = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        stage = [torch.empty((per_batch, T, C), dtype=host.dtype, device=dev) for _ in range(2)]
        in_ready = [torch.cuda.Event() for _ in range(2)]
        in_free = [torch.cuda.Event() for _ in range(2)]
        out_spec = out_audio = None
        pending = []                                        # (device results, event) kept alive until copied back
        n_batches = -(-S // per_batch)

        def upload(b):
            lo, hi = b * per_batch, min(S, (b + 1) * per_batch)
            with torch.cuda.stream(copy_in):
                copy_in.wait_event(in_free[b & 1])
                stage[b & 1][:hi - lo].copy_(host[lo:hi], non_blocking=True)
                in_ready[b & 1].record(copy_in)

        for e in in_free:
            e.record(compute)
        upload(0)
        for b in range(n_batches):
            lo, hi = b * per_batch, min(S, (b + 1) * per_batch)
            if b + 1 < n_batches:
                upload(b + 1)
            compute.wait_event(in_ready[b & 1])
            x = stage[b & 1][:hi - lo]
            lp = self.features.log_power(x, online=True, chunk_size=self.packet_size)
            in_free[b & 1].record(compute)                  # the staging buffer may be refilled once the features exist
            _, spec = self.lda.decode(lp, order=4, step=5, first_row=0, smooth=True, want_labels=False)
            audio = self.gl.synthesize(spec, None, seed + b)
            done = torch.cuda.Event()
            done.record(compute)
            if out_spec is None:
                key = (S,) + tuple(spec.shape[1:]) + tuple(audio.shape[1:])
                if pinned_outputs and getattr(self, '_out_key', None) == key:
                    out_spec, out_audio = self._out_bufs
                else:
                    out_spec = torch.empty((S,) + tuple(spec.shape[1:]), dtype=torch.float64).pin_memory()
                    out_audio = torch.empty((S,) + tuple(audio.shape[1:]), dtype=torch.int16).pin_memory()
                    if pinned_outputs:
                        self._out_key, self._out_bufs = key, (out_spec, out_audio)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                out_spec[lo:hi].copy_(spec, non_blocking=True)
                out_audio[lo:hi].copy_(audio, non_blocking=True)
                spec.record_stream(copy_out)
                audio.record_stream(copy_out)
            pending.append((spec, audio))
        copy_out.synchronize()
        compute.synchronize()
        if pinned_outputs:
            return out_spec.numpy(), out_audio.numpy()
        return out_spec.numpy().copy(), out_audio.numpy().copy()


def session_shard(n_sessions, rank=None, world=None):
    """[lo, hi) of the sessions this rank decodes when a job is spread over the GPUs of one box, one process per GPU:
    sessions are independent end to end, so the decode path has no collective - every rank takes a contiguous, balanced
    slice (sizes differ by at most one).  rank / world default to torch.distributed's, else RANK / WORLD_SIZE, else 0 / 1."""
    import os
    if rank is None or world is None:
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
    if rank is None or world is None:
        rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(n_sessions), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


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
    if n_frames > first:
        noise[first:] = rs.rand(n_frames - first, block)      # one call = the same MT19937 stream as a rand(480) per frame
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


class realtime:
    """Scheduling of the process that runs the node graph, for the closed-loop (latency) path: pins the calling thread to one
    core and asks for SCHED_FIFO, as a real-time deployment would; restores both on exit.  Opt-in, and no cure-all: of the
    3-13 frames per 30 s of real-time feed that take 1.5-2 ms instead of 0.2-0.6 ms (profiles/latency_tail_r02.txt) one run
    lost all of them under this policy (max 0.57 ms) and the next kept them - they follow the idle GPU / host between packets,
    not the scheduler class.  Where the process may not change its policy (no CAP_SYS_NICE) it keeps the pinning and says so
    in `.applied`.

        with decode.realtime():
            sender.start_processing(...)"""

    def __init__(self, core=None, priority=50):
        self.core, self.priority, self.applied = core, priority, {}

    def __enter__(self):
        try:
            self._affinity = os.sched_getaffinity(0)
            cores = sorted(self._affinity)
            core = self.core if self.core is not None else cores[len(cores) // 2]
            os.sched_setaffinity(0, {core})
            self.applied['pinned_to'] = core
        except (AttributeError, OSError) as e:
            self._affinity = None
            self.applied['pinned_to'] = 'refused: %s' % e
        try:
            self._policy = os.sched_getscheduler(0)
            self._param = os.sched_getparam(0)
            os.sched_setscheduler(0, os.SCHED_FIFO, os.sched_param(self.priority))
            self.applied['policy'] = 'SCHED_FIFO %d' % self.priority
        except (AttributeError, OSError) as e:
            self._policy = None
            self.applied['policy'] = 'refused: %s' % e
        return self

    def __exit__(self, *exc):
        if self._policy is not None:
            try:
                os.sched_setscheduler(0, self._policy, self._param)
            except OSError:
                pass
        if self._affinity is not None:
            try:
                os.sched_setaffinity(0, self._affinity)
            except OSError:
                pass
        return False


def _read_datasets(path, names):
    """Datasets `names` of an HDF5 file: through h5py when it is importable, else through the built-in reader of the flat
    layout these files have (sgs/hdf5lite.py); the .npz of the same stem is read when that is what exists (artefacts written
    by round-1 builds of this package)."""
    stem = os.path.splitext(path)[0]
    if os.path.exists(path):
        try:
            import h5py
        except ImportError:
            from sgs import hdf5lite
            return hdf5lite.read(path, names)
        with h5py.File(path, 'r') as hf:
            return {k: hf[k][...] for k in names}
    with np.load(stem + '.npz') as z:
        return {k: z[k] for k in names}


def _write_datasets(path, datasets):
    """HDF5 file with `datasets` in its root group (h5py when importable, else sgs/hdf5lite.py)."""
    try:
        import h5py
    except ImportError:
        from sgs import hdf5lite
        hdf5lite.write(path, datasets)
        return
    with h5py.File(path, 'w') as hf:
        for k, v in datasets.items():
            hf.create_dataset(k, data=v)


def load_params(session_dir):
    """(pickled estimators, medians_array, bad_channels, select) of a training session (decode.py:299-306)."""
    d = _read_datasets(os.path.join(session_dir, 'params.h5'), ('medians_array', 'bad_channels', 'estimators', 'select'))
    return d['estimators'].tobytes(), d['medians_array'], d['bad_channels'], d['select']


def load_seeg(seeg_file):
    """(eeg, sfreq) of a stored recording (decode.py:309-313)."""
    d = _read_datasets(seeg_file, ('sEEG', 'sEEG_sr'))
    return d['sEEG'], int(np.asarray(d['sEEG_sr']).reshape((1,))[0])


# set by the caller before store_decoding_to_file, as the reference's __main__ does (decode.py:259-267, 275)
run_dir = None
config = None


def store_decoding_to_file(spectrogram, output_audio, received_sEEG, sfreq, run_dir=None, config=None):
    """audio.wav (16 kHz int16), sEEG.hdf, spectrogram.npy and decode.ini (decode.py:186-219; the plot is out of scope).
    The reference's signature; like there the target directory and the configuration default to the module globals
    `run_dir` / `config`, and may be passed as keywords instead."""
    from scipy.io.wavfile import write as wavwrite
    target = run_dir if run_dir is not None else globals()['run_dir']
    cfg = config if config is not None else globals()['config']
    if target is None:
        raise ValueError('decode.run_dir is not set: assign it (as decode.py\'s __main__ does) or pass run_dir=...')
    wavwrite(os.path.join(target, 'audio.wav'), 16000, np.asarray(output_audio))
    logger.info('Decoded audio written to {}'.format(os.path.join(target, 'audio.wav')))
    _write_datasets(os.path.join(target, 'sEEG.hdf'), {'sEEG': np.asarray(received_sEEG), 'sEEG_sr': np.int32(sfreq)})
    np.save(os.path.join(target, 'spectrogram.npy'), spectrogram)
    if cfg is not None:
        with open(os.path.join(target, 'decode.ini'), 'w') as configfile:
            cfg.write(configfile)
    logger.info('Decoding artefacts written to {}'.format(target))
