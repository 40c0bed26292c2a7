"""Drops the excluded channel columns from every chunk (reference: livenodes/ChannelSelector.py:10-12)."""
import numpy as np

from . import Node


class ChannelSelector(Node.Node):
    def __init__(self, exclude=None, name='ChannelSelector'):
        super().__init__(name=name)
        self.bad_channels = exclude
        self._keep = None               # column index of the channels that stay, built on the first chunk
        self._width = None

    def add_data(self, data_frame, data_id=0):
        data_frame = np.asarray(data_frame)
        if data_frame.ndim != 2:
            self.output_data(np.delete(data_frame, self.bad_channels, axis=1))
            return
        if self._width != data_frame.shape[1]:
            self._width = data_frame.shape[1]
            self._keep = np.delete(np.arange(self._width), self.bad_channels)       # same index semantics as np.delete
        # np.delete(frame, bad, axis=1) == frame[:, keep]; always a fresh array, as in the reference
        self.output_data(data_frame.copy() if len(self._keep) == self._width else data_frame[:, self._keep])
