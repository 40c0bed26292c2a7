"""Per-mel-bin LDA prediction node (reference: livenodes/LDASynthesis.py:10-28).

`params` is the pickled list of fitted estimators exactly as train.py stores it (anything exposing coef_,
intercept_, classes_); the 40 x 9 x 150 weights are packed once and scored on the device per frame."""
import pickle

import numpy as np

from livenodes import Node
from sgs.lda import LdaDecoder


class LDASynthesis(Node.Node):
    def __init__(self, params, select, name='LDASynthesis'):
        super().__init__(name=name)
        self.estimators = pickle.loads(params)
        self.nb_bins = len(self.estimators)
        self.select = select
        self._dec = None
        self._chain = None              # set by sgs.chain.FusedChain when the feature node above runs the chain

    def add_data(self, frame, data_id=0):
        ch = self._chain
        if ch is not None and frame is ch.cur_rows:
            ch.cur_labels = labels = ch.labels[ch.q].copy()
            self.output_data(labels)
            return
        frame = np.asarray(frame, dtype=np.float64).reshape((1, -1))
        if self._dec is None:
            # medians are irrelevant for label output; a dummy table keeps the fused kernel's interface
            self._dec = LdaDecoder(self.estimators, np.asarray(self.select), np.zeros((self.nb_bins, 9)))
        labels, _ = self._dec.decode(frame, want_spec=False)
        self.output_data(labels[0].copy())
