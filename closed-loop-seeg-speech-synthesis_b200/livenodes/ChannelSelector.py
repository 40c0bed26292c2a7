"""Drops the excluded channel columns from every chunk (reference: livenodes/ChannelSelector.py:10-12)."""
import numpy as np

from . import Node


class ChannelSelector(Node.Node):
    def __init__(self, exclude=None, name='ChannelSelector'):
        super().__init__(name=name)
        self.bad_channels = exclude

    def add_data(self, data_frame, data_id=0):
        self.output_data(np.delete(data_frame, self.bad_channels, axis=1))
