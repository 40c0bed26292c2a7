"""De-quantisation node: medians lookup + gaussian smoothing across bins (reference: livenodes/Dequantization.py)."""
import numpy as np

from livenodes import Node
from sgs import _lib
from sgs.design import gaussian_taps


class Dequantization(Node.Node):
    def __init__(self, medians_array, name='ChannelSelector'):
        super().__init__(name=name)
        self.medians_array = medians_array
        self.c = np.arange(len(self.medians_array))
        self._med = np.ascontiguousarray(medians_array, dtype=np.float64)
        self._taps = np.ascontiguousarray(gaussian_taps(0.5), dtype=np.float64)
        self._chain = None              # set by sgs.chain.FusedChain

    def add_data(self, data_frame, data_id=0):
        ch = self._chain
        if ch is not None and data_frame is ch.cur_labels:
            ch.cur_spec = spec = ch.spec[ch.q].copy()
            self.output_data(spec)
            return
        labels = np.ascontiguousarray(np.asarray(data_frame, dtype=np.float64).reshape(1, -1))
        out = np.empty_like(labels)
        _lib.ensure_init()
        _lib.check(_lib.lib().sgs_dequantize(_lib.ptr(self._med), self._med.shape[0], self._med.shape[1], _lib.ptr(self._taps),
                                             len(self._taps) // 2, _lib.ptr(labels), 1, 1, _lib.ptr(out), None))
        self.output_data(out[0])
