"""Framing node with the reference's constructor (livenodes/FrameBuffer.py:13).

In the reference this class also carries the causal IIR filtering of the feature extractor; that work now
lives in the device stream behind ECogFeatCalc, so this node only frames un-filtered data (host ring buffer,
fractional frame shifts, optional warm-start zero fill) for callers that use it directly."""
import numpy as np

from . import Node


class FrameBuffer(Node.Node):
    def __init__(self, frame_size_ms, frame_shift_ms, sample_rate, filter_coefficients=None, warm_start=False,
                 name="FrameBuffer"):
        super().__init__(name=name)
        if filter_coefficients is not None:
            raise NotImplementedError("filtering FrameBuffers exist only inside ECogFeatCalc, where they run on the "
                                      "device (sgs_feat_stream_push); construct an ECogFeatCalc node instead")
        self.sample_rate = float(sample_rate)
        self.frame_shift_ms = float(frame_shift_ms)
        self.frame_size = int((float(frame_size_ms) / 1000.0) * self.sample_rate)
        self.warm_start = warm_start
        self.total_delay = (self.frame_size / self.sample_rate) * 1000.0
        self.first_frame_at_ms = (float(self.frame_size) / self.sample_rate) * 1000.0
        self.reset_buffer()

    def reset_buffer(self):
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
