"""Sink that collects frames.  Like the reference (livenodes/Receiver.py:16-27) the store is a
multiprocessing.Manager list so that frames appended in a forked feeder process are visible to the parent.

The reference pays one Manager round trip (pickle + socket + unpickle, ~0.1 ms) per frame and receiver, on the
thread that runs the whole graph.  Here frames are kept in a process-local list and moved to the Manager list when
`get_data` / `stop_processing` / `flush` run in the collecting process, when that process exits (a multiprocessing
finalizer, so it also runs in a forked feeder) or is terminated (SIGTERM hook below), and when more than
`max_held_bytes` have piled up.  Nothing is handed over while frames stream, by default: measured on the 128-channel
chain (tools/latency_tail.py, 30 s legs), every frame slower than 1 ms - 9 of 2990 with packets arriving in real time,
a 45 ms stall when a recording is replayed back to back - came from a hand-over, also when it ran on a flusher thread
in small batches (the thread pickles under the interpreter lock the graph thread needs); without hand-overs the slowest
frame took 0.50 ms back to back and 0.72 ms in real time.  A process that polls `get_data` on a Receiver fed by ANOTHER
process therefore sees the frames only after the feeder has flushed; `flush_interval=<seconds>` restores a timed
hand-over (on a flusher thread, at most `max_batch` frames at a time), `flush_interval=0` the reference's per-frame one."""
import multiprocessing
import multiprocessing.util
import os
import queue
import signal
import threading
import time

from . import Node

# Receivers that hold frames in THIS process, and the SIGTERM hook that hands those frames over before a terminated feeder
# dies: Sender.stop_processing (like the reference's, Sender.py:69-70) ends the feeder with Process.terminate(); SIGTERM
# skips multiprocessing's finalizers, so without the hook up to flush_interval of frames would never reach the Manager
# list, while the reference, appending frame by frame, loses nothing.
_live = {'pid': None, 'receivers': [], 'previous': None}


def _on_sigterm(signum, frame):
    for r in list(_live['receivers']):
        try:
            r.flush()
        except Exception:
            pass
    prev = _live['previous']
    if callable(prev):
        prev(signum, frame)
        return
    signal.signal(signal.SIGTERM, signal.SIG_DFL)
    os.kill(os.getpid(), signal.SIGTERM)


def _register(receiver):
    pid = os.getpid()
    if _live['pid'] != pid:
        _live.update(pid=pid, receivers=[], previous=None)
        # only a forked feeder is ended by terminate(); the collecting parent flushes in get_data / stop_processing
        if multiprocessing.parent_process() is not None and threading.current_thread() is threading.main_thread():
            try:
                prev = signal.signal(signal.SIGTERM, _on_sigterm)
                _live['previous'] = prev if prev not in (signal.SIG_DFL, signal.SIG_IGN, None, _on_sigterm) else None
            except (ValueError, OSError):
                pass
    _live['receivers'].append(receiver)


class Receiver(Node.Node):
    def __init__(self, perform_timing=False, dont_time=False, name='Receiver', flush_interval=None, max_batch=32,
                 max_held_bytes=2 << 30):
        super().__init__(has_outputs=False, dont_time=dont_time, name=name)
        self._manager = multiprocessing.Manager()
        self.data = self._manager.list([])
        self.perform_timing = perform_timing
        env = os.environ.get('SGS_RECEIVER_FLUSH_INTERVAL')
        self.flush_interval = (None if env in ('', 'none', 'None') else float(env)) if env is not None else flush_interval
        self.max_batch = int(os.environ.get('SGS_RECEIVER_MAX_BATCH', max_batch))
        self.max_held_bytes = max_held_bytes
        self._held_bytes = 0
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
        _register(self)

    def _flusher(self, q):
        while True:
            batch = q.get()
            try:
                # batches that failed earlier go first, so that frames reach the list in the order they were produced
                self._unsent.append(batch)
                while self._unsent:
                    self.data.extend(self._unsent[0])
                    self._unsent.pop(0)
            except Exception:                # keep the frames, in order; the next hand-over (or synchronous flush) retries
                pass
            finally:
                q.task_done()

    def add_data(self, sample, data_id=None):
        if self._local_pid != os.getpid():
            self._adopt_process()
        now = time.time()
        self._local.append([now, sample] if self.perform_timing else sample)
        self._held_bytes += getattr(sample, 'nbytes', 64)
        if self.flush_interval is None:
            if self._held_bytes > self.max_held_bytes:
                self.flush()
            return
        if now - self._last_flush >= self.flush_interval or (self.max_batch > 0 and len(self._local) >= self.max_batch):
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
            self._held_bytes = 0
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
            self._held_bytes = 0
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
