"""Streaming Griffin-Lim synthesis node (reference: livenodes/GriffinLim.py:13-174).

Each add_data pushes one log-mel frame to the device stream (sgs_gl_node_push), which synthesises that frame's
480-sample block, overlap-adds it with the blocks still in its ring, applies the output low-pass with carried
state and returns the int16 hop.  The initial waveform of every block is drawn with np.random.rand(480) on the
host in the reference's order, so a seeded run reproduces the reference's audio; pass device_noise=True to draw
on the device instead."""
import time

import numpy as np

from . import Node
from sgs import _lib
from sgs.griffinlim import GriffinLimNodeOp


class GriffinLimSynthesis(Node.Node):
    def __init__(self, originalFrameSizeMs, frameShiftMs, sampleRate, melCoeffCount, numReconstructionIterations=5,
                 extraContext=0, cutoff=7900, normFactor=1.0, useLogMels=True, name='GriffinLim', device_noise=False):
        super().__init__(name=name)
        if extraContext != 0:
            raise NotImplementedError("only extraContext=0 (decode.py:162-164) is implemented: a longer block changes the "
                                      "register layout of the block kernel")
        self._op = GriffinLimNodeOp(originalFrameSizeMs, frameShiftMs, sampleRate, melCoeffCount,
                                    numReconstructionIterations, cutoff, normFactor, use_log_mels=useLogMels)
        plan = self._op.plan
        self.useLogMels = useLogMels
        self.frameShiftMs = float(frameShiftMs)
        self.sampleRate = float(sampleRate)
        self.fftSize = plan.fft_size
        self.frameShift = plan.hop
        self.contextWidth = plan.context_width
        self.blockLen = plan.block_len
        self.normFactor = normFactor
        self.numReconstructionIterations = numReconstructionIterations
        self.fftWindow = plan.window
        self.filterNumerator, self.filterDenominator = plan.lp_b, plan.lp_a
        self.outputBufferPosMs = 0
        self.framePos = 0
        self.rfc = 0
        self.startTime = time.time()
        self.device_noise = device_noise
        self._pcm = np.empty(16 * 192, dtype=np.int16)
        self._chain = None              # set by sgs.chain.FusedChain
        self._pos_base = 0

    _REBASE_AT = 1 << 30

    def _head(self):
        """Write head before the next frame, in the 32-bit coordinates the library is given: the absolute sample count (the
        reference's float expression, quirk Q7) minus a base that moves up before the count outgrows an int32 - the library
        only uses differences of positions (the reference keeps them modulo its ring length and so runs for ever too)."""
        prev = int((self.outputBufferPosMs / 1000.0) * self.sampleRate)
        if prev - self._pos_base >= self._REBASE_AT:
            delta = prev - self._pos_base
            _lib.check(_lib.lib().sgs_gl_node_rebase(self._op.handle(), delta))
            self._pos_base += delta
        return prev - self._pos_base

    def _reserve(self, n):
        """Host bookkeeping of add_data for the next n frames at once (fused chain): write-head positions, the
        np.random.rand(480) draws in frame order, and which frames emit audio."""
        prev = self._head()
        pos = np.empty(n, dtype=np.int32)
        emit = [False] * n
        noise = None if self.device_noise else np.zeros((n, self.blockLen * self.frameShift))
        for i in range(n):
            self.framePos += 1
            self.outputBufferPosMs += self.frameShiftMs
            pos[i] = int((self.outputBufferPosMs / 1000.0) * self.sampleRate) - self._pos_base
            emit[i] = not (self.framePos < self.blockLen - self.contextWidth)
            if emit[i] and noise is not None:
                noise[i] = np.random.rand(self.blockLen * self.frameShift)
        return pos, prev, noise, emit

    def add_data(self, dataFrame, data_id=0):
        ch = self._chain
        if ch is not None and dataFrame is ch.cur_spec:
            q = ch.q
            if not ch.emit[q]:
                return np.array([])
            self.rfc += ch.pcm_hi[q] - ch.pcm_lo[q]
            self.output_data(ch.pcm[ch.pcm_lo[q]:ch.pcm_hi[q]].copy())
            return
        frame = np.ascontiguousarray(np.asarray(dataFrame, dtype=np.float64).reshape(1, -1))
        self.framePos += 1
        prev = self._head()
        self.outputBufferPosMs += self.frameShiftMs
        pos = np.array([int((self.outputBufferPosMs / 1000.0) * self.sampleRate) - self._pos_base], dtype=np.int32)
        first = self.framePos < self.blockLen - self.contextWidth
        noise = None
        if not first and not self.device_noise:
            noise = np.random.rand(self.blockLen * self.frameShift)          # GriffinLim.py:90, same draw order
        n_pcm = _lib.c_int(0)
        _lib.check(_lib.lib().sgs_gl_node_push(self._op.handle(), _lib.ptr(frame), 1, _lib.ptr(pos), prev, _lib.ptr(noise),
                                               self.framePos, _lib.ptr(self._pcm), _lib.C.byref(n_pcm), None))
        if first:
            return np.array([])
        self.rfc += n_pcm.value
        self.output_data(self._pcm[:n_pcm.value].copy())
