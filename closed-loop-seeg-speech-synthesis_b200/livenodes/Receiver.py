"""Sink that collects frames.  Like the reference (livenodes/Receiver.py:16-27) the store is a
multiprocessing.Manager list so that frames appended in a forked feeder process are visible to the parent."""
import multiprocessing
import time

from . import Node


class Receiver(Node.Node):
    def __init__(self, perform_timing=False, dont_time=False, name='Receiver'):
        super().__init__(has_outputs=False, dont_time=dont_time, name=name)
        self._manager = multiprocessing.Manager()
        self.data = self._manager.list([])
        self.perform_timing = perform_timing

    def add_data(self, sample, data_id=None):
        self.data.append([time.time(), sample] if self.perform_timing else sample)

    def get_data(self, clear=False):
        out = list(self.data)
        if clear:
            self.data[:] = []
        return out
