"""Mel filter bank with the reference's matrices and method names (local/MelFilterBank.py:5-83).

Table construction is host-side (O(spec_size x bands), once per configuration); the device kernels use the
2-tap form of the inverse (sgs.design.MelTables.inv_idx / inv_w).  The array methods below are small dense
products kept for API compatibility with callers that hold a MelFilterBank object."""
import numpy as np

from sgs.design import MelTables


class MelFilterBank:
    def __init__(self, specSize, numCoefficients, sampleRate):
        self._tables = MelTables(specSize, numCoefficients, sampleRate)
        self.melMatrix = self._tables.mel
        self.melInvMatrix = self._tables.inv

    @staticmethod
    def makeNormal(x):
        x[np.isnan(x)] = 0
        x[np.isinf(x)] = 0
        return x

    @staticmethod
    def fuzz(x):
        return x + 0.0000001

    def toMelScale(self, spectrogram):
        return np.dot(spectrogram, self.melMatrix)

    def fromMelScale(self, melSpectrogram):
        return np.dot(melSpectrogram, self.melInvMatrix)

    toMels = toMelScale
    fromMels = fromMelScale

    def toLogMels(self, spectrogram):
        return self.makeNormal(np.log(self.fuzz(self.toMelScale(spectrogram))))

    def fromLogMels(self, melSpectrogram):
        return self.makeNormal(self.fromMelScale(np.exp(melSpectrogram)))
