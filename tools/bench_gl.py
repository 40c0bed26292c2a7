"""Device time of the node-semantics Griffin-Lim kernels alone (k_gl_blocks dominates): S sessions x T frames.
Usage: python tools/bench_gl.py [sessions] [frames]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
from sgs import _lib  # noqa: E402
from sgs.griffinlim import GriffinLimNodeOp  # noqa: E402

if __name__ == '__main__':
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    _lib.ensure_init(0)
    op = GriffinLimNodeOp(16, 10, 16000, 40, 8, 7900, 10)
    torch.manual_seed(0)
    spec = (torch.randn((S, T, 40), dtype=torch.float64, device='cuda') * 1.5 - 2.0)
    for _ in range(2):
        pcm = op.synthesize(spec, None, 3)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    reps = 5
    for _ in range(reps):
        pcm = op.synthesize(spec, None, 3)
    torch.cuda.synchronize()
    out = {"sessions": S, "frames": T}
    for k in ('gl_blocks', 'gl_ola', 'lowpass'):
        ms, n = _lib.profile_read(k)
        out[k + "_ms"] = ms / max(n, 1)
    out["blocks_per_s"] = S * (T - 1) / (out["gl_blocks_ms"] * 1e-3)
    out["checksum"] = int(pcm.to(torch.int64).sum().item())
    print(json.dumps(out))
