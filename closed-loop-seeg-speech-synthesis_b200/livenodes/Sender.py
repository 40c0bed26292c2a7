"""Streams an in-memory array chunk by chunk from a forked feeder process (reference: livenodes/Sender.py).

The whole downstream graph runs inside that child process, which is why every CUDA-backed node creates its
device state lazily on its first add_data (sgs._lib.ensure_init is per process)."""
import gc
import time
from multiprocessing import Process

import numpy as np

from . import Node


class Sender(Node.Node):
    def __init__(self, data, sample_rate, frame_size_ms, asap=False, name='Sender'):
        super().__init__(has_inputs=False, name=name)
        self.data = data
        self.sample_rate = sample_rate
        self.frame_size_ms = frame_size_ms
        self.asap = asap
        self.feeder_process = None

    def sender_process(self):
        step = int(self.sample_rate / 1000 * self.frame_size_ms)
        period = (1.0 / 1000.0) * self.frame_size_ms
        t_start = t_ref = time.time()
        for first in range(0, len(self.data), step):
            if not self.asap:
                while time.time() - t_ref < period:
                    time.sleep(0.000001)
                t_ref = t_start + first / self.sample_rate
            self.output_data(np.array(self.data[first:first + step]))

    def _spawn(self):
        self.feeder_process = Process(target=self.sender_process)
        self.feeder_process.start()

    def send_new(self, data):
        if self.feeder_process is None:
            gc.collect()
            self.data = data
            self._spawn()
        super().start_processing()

    def wait_for_completion(self):
        if self.feeder_process is not None:
            self.feeder_process.join()
        self.feeder_process = None

    def start_processing(self, recurse=True):
        if self.feeder_process is None and self.data is not None:
            self._spawn()
        super().start_processing(recurse)

    def stop_processing(self, recurse=True):
        super().stop_processing(recurse)
        if self.feeder_process is not None:
            self.feeder_process.terminate()
            # the feeder's receivers hand their last batch over on SIGTERM (Receiver.py): wait for that before the caller reads them
            self.feeder_process.join(5.0)
        self.feeder_process = None
