"""Framing node with the reference's constructor (livenodes/FrameBuffer.py:13-177).

Frames a stream of chunks (host ring of the samples still needed, fractional frame shifts rounded exactly as
FrameBuffer.py:177, optional warm-start zero fill).  With `filter_coefficients` (second-order sections) every chunk
first goes through the causal IIR cascade on the device (sgs_sos_stream_push): the filter state stays resident between
chunks, starts from sosfilt_zi scaled by the first sample (cold start, FrameBuffer.py:90-92) or un-scaled with the
zero fill pushed through the filter first (warm start, FrameBuffer.py:95-98).  The three filtering FrameBuffers inside
ECogFeatCalc do not use this class: that sub-graph is one fused device stream (sgs_feat_stream_push).
Q5 of SURVEY.md: the reference overwrites history it still needs when a single chunk exceeds its 2048-sample ring;
chunk-size-independent behaviour (that of small chunks) is what this node implements for all chunk sizes."""
import numpy as np

from . import Node

MAX_PUSH = 4096     # samples per device push


class FrameBuffer(Node.Node):
    def __init__(self, frame_size_ms, frame_shift_ms, sample_rate, filter_coefficients=None, warm_start=False,
                 name="FrameBuffer"):
        super().__init__(name=name)
        self.filter_coefficients = None if filter_coefficients is None else np.ascontiguousarray(filter_coefficients, dtype=np.float64)
        if self.filter_coefficients is not None and (self.filter_coefficients.ndim != 2 or self.filter_coefficients.shape[1] != 6):
            raise ValueError("filter_coefficients must be second-order sections (n_sections x 6)")
        self._filter = None             # device stream, created on the first chunk in the process that pushes data
        self.sample_rate = float(sample_rate)
        self.frame_shift_ms = float(frame_shift_ms)
        self.frame_size = int((float(frame_size_ms) / 1000.0) * self.sample_rate)
        self.warm_start = warm_start
        self.total_delay = (self.frame_size / self.sample_rate) * 1000.0
        self.first_frame_at_ms = (float(self.frame_size) / self.sample_rate) * 1000.0
        self.reset_buffer()

    def _close_filter(self):
        if getattr(self, '_filter', None) is not None:
            from sgs import _lib
            _lib.lib().sgs_sos_stream_destroy(self._filter)
        self._filter = None

    def __del__(self):
        try:
            self._close_filter()
        except Exception:
            pass

    def _filtered(self, data):
        """scipy.signal.sosfilt(filter_coefficients, data, axis=0, zi=carried state) on the device."""
        from scipy.signal import sosfilt_zi
        from sgs import _lib
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        if self._filter is None:
            _lib.ensure_init()
            sos = self.filter_coefficients
            zi = np.ascontiguousarray(sosfilt_zi(sos), dtype=np.float64)
            h = _lib.c_void_p()
            _lib.check(_lib.lib().sgs_sos_stream_create(_lib.C.byref(h), _lib.ptr(sos), _lib.ptr(zi), sos.shape[0], data.shape[1],
                                                        int(bool(self.warm_start))))
            self._filter = h
        out = np.empty(data.shape, dtype=np.float64)
        for lo in range(0, len(data), MAX_PUSH):
            block = np.ascontiguousarray(data[lo:lo + MAX_PUSH])
            y = out[lo:lo + len(block)]
            _lib.check(_lib.lib().sgs_sos_stream_push(self._filter, _lib.ptr(block), int(block.dtype == np.float64), len(block),
                                                      _lib.ptr(y), None))
        return out

    def reset_buffer(self):
        self._close_filter()
        self._hist = None
        self._seen = 0
        self.frame_count = 0
        self.next_frame_at = self.frame_size

    def add_data(self, data, data_id=None):
        data = np.asarray(data)
        if data.ndim == 1:
            data = data.reshape(-1, 1)
        if self._hist is None:
            self._hist = np.zeros((0, data.shape[1]))
            if self.warm_start:
                fill = self.frame_size - int((self.frame_shift_ms / 1000.0) * self.sample_rate)
                assert fill > 0
                self.add_data(np.zeros((fill, data.shape[1])))
        if self.filter_coefficients is not None and len(data):
            data = self._filtered(data)
        self._hist = np.vstack([self._hist, data])
        self._seen += len(data)
        base = self._seen - len(self._hist)
        while self.next_frame_at <= self._seen:
            end = self.next_frame_at - base
            self.frame_count += 1
            self.next_frame_at = round(((self.first_frame_at_ms + self.frame_count * self.frame_shift_ms) / 1000.0)
                                       * self.sample_rate)
            self.output_data(self._hist[end - self.frame_size:end].copy())
        keep = self._seen - (self.next_frame_at - self.frame_size)
        if keep < len(self._hist):
            self._hist = self._hist[len(self._hist) - max(keep, 0):]
