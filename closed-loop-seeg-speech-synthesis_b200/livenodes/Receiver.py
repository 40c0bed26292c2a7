"""Sink that collects frames.  Like the reference (livenodes/Receiver.py:16-27) the store is a
multiprocessing.Manager list so that frames appended in a forked feeder process are visible to the parent.

The reference pays one Manager round trip (pickle + socket + unpickle, ~0.1 ms) per frame and receiver, on the
thread that runs the whole graph.  Here frames are kept in a process-local list and moved to the Manager list in
batches: every `flush_interval` seconds of streaming, when the process exits (a multiprocessing finalizer, so it
also runs in a forked feeder), and whenever `get_data` / `stop_processing` run in the collecting process.
The periodic hand-over runs on a flusher thread of the collecting process, not on the graph thread: a batch of
0.25 s of 128-channel sEEG is a 0.3-0.7 ms Manager round trip, which on the graph thread was the p99 of the frame
latency when packets arrive in real time (one packet in eight paid it).  `flush_interval=0` restores the per-frame
behaviour."""
import multiprocessing
import multiprocessing.util
import os
import queue
import threading
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
        self._queue = None
        self._thread = None
        self._unsent = []

    def _adopt_process(self):
        """First frame seen in this process (possibly a forked child): start a fresh local batch and make sure it is
        handed over before the process goes away."""
        self._local = []
        self._local_pid = os.getpid()
        self._last_flush = time.time()
        self._queue = None                  # threads do not survive a fork: the flusher starts with the first timed hand-over
        self._thread = None
        self._unsent = []
        multiprocessing.util.Finalize(self, self.flush, exitpriority=100)

    def _flusher(self, q):
        while True:
            batch = q.get()
            try:
                self.data.extend(batch)
            except Exception:                # keep the frames; the next synchronous flush retries and reports
                self._unsent.append(batch)
            finally:
                q.task_done()

    def add_data(self, sample, data_id=None):
        if self._local_pid != os.getpid():
            self._adopt_process()
        now = time.time()
        self._local.append([now, sample] if self.perform_timing else sample)
        if now - self._last_flush >= self.flush_interval:
            if self.flush_interval <= 0:
                self.flush()
                return
            # timed hand-over: off the graph thread
            if self._thread is None:
                self._queue = queue.Queue()
                self._thread = threading.Thread(target=self._flusher, args=(self._queue,), daemon=True,
                                                name=self.name + '-flusher')
                self._thread.start()
            batch, self._local = self._local, []
            self._queue.put(batch)
            self._last_flush = now

    def flush(self):
        """Synchronous hand-over of everything this process still holds (batches queued for the flusher first)."""
        if self._local_pid == os.getpid():
            if self._thread is not None:
                self._queue.join()
            pending, self._unsent = self._unsent, []
            if self._local:
                pending.append(self._local)
                self._local = []
            for batch in pending:
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
