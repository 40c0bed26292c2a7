"""Fused streaming chain: ECogFeatCalc -> LDASynthesis -> Dequantization -> GriffinLimSynthesis in one device
round trip per packet (sgs_chain_push).

The reference runs these four nodes synchronously and depth-first, one Python callback per node and 10 ms frame
(decode.py:152-183 wires them; Node.py:158-164 pushes).  When a feature node finds exactly that wiring below
itself, it computes every product of the chain for all frames a packet completes in a single C call and the three
downstream nodes emit their share of it when the frame reaches them - the callbacks, their order and the arrays
they deliver are unchanged, only the device work is hoisted to the front of the packet.  A frame is recognised by
object identity (`frame is chain.cur_rows`): anything else reaching those nodes takes their ordinary path."""
import os

import numpy as np

from . import _lib
from .lda import LdaDecoder

MAX_FRAMES = 16
BLOCK = 480


def find_chain(feat_node):
    """(lda_node, deq_node, gl_node) when the graph below feat_node is the reference's decode wiring, else None."""
    if os.environ.get('SGS_FUSED_CHAIN', '1') == '0':
        return None
    from livenodes import LDASynthesis, Dequantization, GriffinLim

    def only(node, cls):
        hits = [n for n in node.output_classes if isinstance(n, cls)]
        return hits[0] if len(hits) == 1 else None

    lda = only(feat_node, LDASynthesis.LDASynthesis)
    deq = only(lda, Dequantization.Dequantization) if lda is not None else None
    gl = only(deq, GriffinLim.GriffinLimSynthesis) if deq is not None else None
    if gl is None:
        return None
    # each of the three must be fed by its predecessor alone, or hoisting the work would reorder it
    for node, src in ((lda, feat_node), (deq, lda), (gl, deq)):
        if len(node.input_classes) != 1 or node.input_classes[0] is not src:
            return None
    return lda, deq, gl


class FusedChain:
    def __init__(self, feat_node, lda_node, deq_node, gl_node):
        self.feat_node, self.lda_node, self.deq_node, self.gl_node = feat_node, lda_node, deq_node, gl_node
        self.lda = LdaDecoder(lda_node.estimators, np.asarray(lda_node.select), deq_node._med)
        width = feat_node._n_channels * (feat_node.model_order + 1)
        h = _lib.c_void_p()
        _lib.check(_lib.lib().sgs_chain_create(_lib.C.byref(h), feat_node._stream, feat_node._n_channels, self.lda.handle(),
                                               gl_node._op.handle()))
        self._h = h
        nb = self.lda.n_bins
        self.rows = np.empty((MAX_FRAMES, width), dtype=np.float64)
        self.labels = np.empty((MAX_FRAMES, nb), dtype=np.float64)
        self.spec = np.empty((MAX_FRAMES, nb), dtype=np.float64)
        self.pcm = np.empty(MAX_FRAMES * 192, dtype=np.int16)
        self._p_rows, self._p_labels, self._p_spec, self._p_pcm = (_lib.ptr(a) for a in (self.rows, self.labels, self.spec, self.pcm))
        self._n_pcm = _lib.c_int(0)
        self._p_n_pcm = _lib.C.byref(self._n_pcm)
        self._push = _lib.lib().sgs_chain_push
        # per-push frame table: emit flag and pcm slice of every frame
        self.q = 0
        self.emit = [False] * MAX_FRAMES
        self.pcm_lo = [0] * MAX_FRAMES
        self.pcm_hi = [0] * MAX_FRAMES
        self.cur_rows = self.cur_labels = self.cur_spec = None
        lda_node._chain = deq_node._chain = gl_node._chain = self

    def close(self):
        for node in (self.lda_node, self.deq_node, self.gl_node):
            if getattr(node, '_chain', None) is self:
                node._chain = None
        if self._h is not None:
            _lib.lib().sgs_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, block, ends, idx):
        """block: (n, C) contiguous samples; ends / idx: int64 arrays of the frames this push completes."""
        nf = len(ends)
        n = len(block)
        if nf == 0:
            _lib.check(self._push(self._h, _lib.ptr(block), int(block.dtype == np.float64), n, None, None, 0, None, 0, None, 0,
                                  None, None, None, None, None, None))
            return 0
        pos, prev, noise, emit = self.gl_node._reserve(nf)
        _lib.check(self._push(self._h, _lib.ptr(block), int(block.dtype == np.float64), n, _lib.ptr(ends), _lib.ptr(idx), nf,
                              _lib.ptr(pos), prev, _lib.ptr(noise), int(self.gl_node.framePos), self._p_rows, self._p_labels,
                              self._p_spec, self._p_pcm, self._p_n_pcm, None))
        off = 0
        for q in range(nf):
            self.emit[q] = emit[q]
            self.pcm_lo[q] = off
            if emit[q]:
                off += int(pos[q]) - prev
            self.pcm_hi[q] = off
            prev = int(pos[q])
        assert off == self._n_pcm.value
        return nf
