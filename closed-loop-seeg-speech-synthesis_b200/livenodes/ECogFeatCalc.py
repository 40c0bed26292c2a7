"""Streaming high-gamma feature node with the reference's constructor and callback contract
(livenodes/ECogFeatCalc.py:19-144).

The reference wires three filtering FrameBuffers, a log-power LambdaNode, a 21-row stack FrameBuffer and a
stacking LambdaNode behind this facade.  Here the whole sub-graph is one stateful device stream
(sgs_feat_stream_push): filter states, the squared-signal history and the stack buffer stay in HBM between
chunks; the host only keeps the fractional frame schedule of FrameBuffer.py:177.  Output per 10 ms frame:
a 1-D float64 array of (model_order+1)*channels values in the reference's c*5+tap order."""
import logging

import numpy as np

from . import Node
from sgs import _lib
from sgs.features import FeatureExtractor

logger = logging.getLogger('ECoGFeatCalc.py')
MAX_PUSH = 128      # samples per device push (csrc/stream.cu)
MAX_FRAMES = 16     # frames one device push can complete (csrc/kernels.cuh:kMaxFramesPerPush)


class ECogFeatCalc(Node.Node):
    def __init__(self, sample_rate, frame_len_ms, frame_shift_ms, model_order=4, step_size=5,
                 line_noise=50, warm_start=True, chunk_size=32, has_inputs=True, name='ECogFeatCalc', fuse_chain=True):
        super().__init__(name=name, has_inputs=has_inputs)
        # warm_start=False (no entry point of the reference passes it): the last filter starts cold, frames are cut from the
        # first sample on, and the stack buffer emits only once it holds model_order * step_size + 1 rows (FrameBuffer.py:91-98)
        self.warm_start = bool(warm_start)
        self.sample_rate = sample_rate
        self.frame_len_ms = frame_len_ms
        self.frame_shift_ms = frame_shift_ms
        self.model_order = model_order
        self.step_size = step_size
        self.chunk_size = int(chunk_size)
        logger.info('Framelength in ms: ' + str(frame_len_ms))
        logger.info('Frameshift in ms: ' + str(frame_shift_ms))
        logger.info('Samplerate: ' + str(sample_rate))
        self._fe = FeatureExtractor(sample_rate, line_noise=line_noise, model_order=model_order, step_size=step_size,
                                    frame_len_ms=frame_len_ms, frame_shift_ms=frame_shift_ms)
        plan = self._fe.plan
        self.high_gamma_filter = plan.filters[0]
        self.first_harmonic_filter = plan.filters[1]
        self.second_harmonic_filter = plan.filters[2] if plan.n_filters == 3 else None
        self.fuse_chain = fuse_chain    # run LDASynthesis/Dequantization/GriffinLim below this node in the same device call
        self._chain = None
        self._stream = None
        self._pending = None            # samples not yet forming a complete chunk_size block (FrameBuffer semantics)
        self._n_channels = None
        self._consumed = 0              # samples handed to the device
        self._frame_count = 0
        sr = float(sample_rate)
        self._first_ms = (float(plan.frame_size) / sr) * 1000.0             # FrameBuffer.py:35
        self._next_end = plan.frame_size                                    # in zero-fill-prefixed coordinates

    # -- device stream, created lazily in the process that pushes data (the graph may run in a forked child) --
    def _open(self, n_channels):
        h = _lib.c_void_p()
        plan = self._fe.plan
        _lib.check(_lib.lib().sgs_feat_stream_create(_lib.C.byref(h), self._fe.handle(), n_channels, plan.frame_size,
                                                     self.model_order, self.step_size))
        if not self.warm_start:
            _lib.check(_lib.lib().sgs_feat_stream_set_cold_start(h, 1))
        self._stream = h
        self._n_channels = n_channels
        self._out = np.empty((16, n_channels * (self.model_order + 1)), dtype=np.float64)
        if self.fuse_chain and self.warm_start:          # the fused chain emits every frame; the cold stack buffer withholds the first ones
            from sgs import chain
            nodes = chain.find_chain(self)
            if nodes is not None:
                self._chain = chain.FusedChain(self, *nodes)

    def reset_buffer(self):
        """Forget all streaming state (the reference re-arms FrameBuffer.reset_buffer per input process)."""
        if self._chain is not None:
            self._chain.close()
            self._chain = None
        if self._stream is not None:
            _lib.lib().sgs_feat_stream_destroy(self._stream)
        self._stream = None
        self._pending = None
        self._consumed = 0
        self._frame_count = 0
        self._next_end = self._fe.plan.frame_size

    def __del__(self):
        try:
            if self._chain is not None:
                self._chain.close()
            if self._stream is not None:
                _lib.lib().sgs_feat_stream_destroy(self._stream)
        except Exception:
            pass

    def _end_of_frame(self, k):
        """End (exclusive, zero-fill-prefixed coordinates) of frame k - FrameBuffer.py:35,177, Python's banker's round."""
        return round(((self._first_ms + k * float(self.frame_shift_ms)) / 1000.0) * float(self.sample_rate))

    def _schedule(self, n_max):
        """(n, ends, idx): how many of the next n_max samples go into this push, and the frames whose end falls inside them
        (FrameBuffer.py:147-177).  A device push completes at most MAX_FRAMES frames: when more would end inside n_max samples
        (sample rates below ~800 Hz at a 10 ms shift, or a short frame_shift_ms) the push is cut just before the end of the
        first frame that does not fit, so no frame is ever deferred past the samples that complete it."""
        zero_fill = self._fe.plan.zero_fill if self.warm_start else 0
        base = zero_fill + self._consumed
        n, ends, idx = n_max, [], []
        while self._next_end <= base + n:
            if len(ends) == MAX_FRAMES:
                n = self._next_end - 1 - base
                break
            ends.append(self._next_end - zero_fill)
            idx.append(self._frame_count)
            self._frame_count += 1
            self._next_end = self._end_of_frame(self._frame_count)
        if n < 1 or (ends and ends[-1] + zero_fill > base + n):
            raise ValueError("frame_shift_ms=%r at sample_rate=%r completes more than %d frames per sample: unsupported"
                             % (self.frame_shift_ms, self.sample_rate, MAX_FRAMES))
        return n, ends, idx

    def add_data(self, data, data_id=None):
        data = np.asarray(data)
        if data.ndim == 1:
            data = data.reshape(-1, 1)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        if self._stream is None:
            self._open(data.shape[1])
        if self._pending is not None and len(self._pending):
            data = np.vstack([self._pending, data])
        usable = (len(data) // self.chunk_size) * self.chunk_size
        self._pending = data[usable:].copy() if usable < len(data) else None
        pos = 0
        while pos < usable:
            n, ends, idx = self._schedule(min(MAX_PUSH, usable - pos))
            block = np.ascontiguousarray(data[pos:pos + n])
            e = np.asarray(ends, dtype=np.int64)
            k = np.asarray(idx, dtype=np.int64)
            ch = self._chain
            if ch is not None:
                # the whole chain below this node runs in this one call; the nodes emit their share frame by frame
                ch.push(block, e, k)
                self._consumed += n
                pos += n
                for q in range(len(ends)):
                    ch.q = q
                    ch.cur_rows = row = ch.rows[q].copy()
                    self.output_data(row)
                ch.cur_rows = ch.cur_labels = ch.cur_spec = None
                continue
            _lib.check(_lib.lib().sgs_feat_stream_push(self._stream, _lib.ptr(block), int(block.dtype == np.float64), n,
                                                       _lib.ptr(e) if len(ends) else None, _lib.ptr(k) if len(ends) else None,
                                                       len(ends), _lib.ptr(self._out), None))
            self._consumed += n
            pos += n
            for q in range(len(ends)):
                if self.warm_start or idx[q] >= self.model_order * self.step_size:
                    self.output_data(self._out[q].copy())
