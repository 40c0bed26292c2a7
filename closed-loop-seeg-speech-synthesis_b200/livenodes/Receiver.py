"""Sink that collects frames.  Like the reference (livenodes/Receiver.py:16-27) the store is a
multiprocessing.Manager list so that frames appended in a forked feeder process are visible to the parent.

The reference pays one Manager round trip (pickle + socket + unpickle, ~0.1 ms) per frame and receiver, on the
thread that runs the whole graph.  Here frames are kept in a process-local list and moved to the Manager list in
batches: every `flush_interval` seconds of streaming, when the process exits (a multiprocessing finalizer, so it
also runs in a forked feeder), and whenever `get_data` / `stop_processing` run in the collecting process.
`flush_interval=0` restores the per-frame behaviour."""
import multiprocessing
import multiprocessing.util
import os
import time

from . import Node


class Receiver(Node.Node):
    def __init__(self, perform_timing=False, dont_time=False, name='Receiver', flush_interval=0.25):
        super().__init__(has_outputs=False, dont_time=dont_time, name=name)
        self._manager = multiprocessing.Manager()
        self.data = self._manager.list([])
        self.perform_timing = perform_timing
        self.flush_interval = flush_interval
        self._local = []
        self._local_pid = None
        self._last_flush = 0.0

    def _adopt_process(self):
        """First frame seen in this process (possibly a forked child): start a fresh local batch and make sure it is
        handed over before the process goes away."""
        self._local = []
        self._local_pid = os.getpid()
        self._last_flush = time.time()
        multiprocessing.util.Finalize(self, self.flush, exitpriority=100)

    def add_data(self, sample, data_id=None):
        if self._local_pid != os.getpid():
            self._adopt_process()
        now = time.time()
        self._local.append([now, sample] if self.perform_timing else sample)
        if now - self._last_flush >= self.flush_interval:
            self.flush()

    def flush(self):
        if self._local_pid == os.getpid() and self._local:
            batch, self._local = self._local, []
            self.data.extend(batch)
        self._last_flush = time.time()

    def stop_processing(self, recurse=True):
        self.flush()
        super().stop_processing(recurse)

    def get_data(self, clear=False):
        self.flush()
        out = list(self.data)
        if clear:
            self.data[:] = []
        return out
