"""Device time of the LDA decode (tensor-core filter + exact fp64 re-scoring + dequantisation) on un-stacked log-power features
resident in HBM: S sessions x W windows x 128 channels.  Usage: python tools/bench_lda.py [sessions] [windows]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
import bench  # noqa: E402
from sgs import _lib  # noqa: E402
from sgs.lda import LdaDecoder  # noqa: E402

if __name__ == '__main__':
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    _lib.ensure_init(0)
    rng = np.random.default_rng(7)
    model, select, medians = bench.trained_model()
    dec = LdaDecoder(model, select, medians)
    torch.manual_seed(0)
    lp = torch.randn((S, W, 128), dtype=torch.float64, device='cuda') * 0.6 + 8.0
    for _ in range(2):
        dec.decode(lp, order=4, step=5, first_row=0, smooth=True, want_labels=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.profile_enable(True)
    reps = 5
    a.record()
    for _ in range(reps):
        dec.decode(lp, order=4, step=5, first_row=0, smooth=True, want_labels=False)
    b.record()
    torch.cuda.synchronize()
    tc_ms, n = _lib.profile_read('lda_tc')
    flop = 2.0 * S * W * 160 * 384 * 3                                     # hi.hi + hi.lo + lo.hi
    print(json.dumps({"sessions": S, "windows": W, "decode_ms": a.elapsed_time(b) / reps, "pack_plus_tc_ms": tc_ms / n,
                      "tf32_tflops_issued": flop / (tc_ms / n * 1e-3) / 1e12, "frames_rescored": dec.last_rescored()}))
