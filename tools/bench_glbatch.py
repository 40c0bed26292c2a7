"""BASELINE config 4: batched offline Griffin-Lim (local/offline.py:131-192 semantics, 800-point frames) of U utterances x
T frames, N iterations, resident in HBM.  Usage: python tools/bench_glbatch.py [utterances] [frames] [iterations]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
from sgs import _lib  # noqa: E402
from sgs.griffinlim import griffin_lim_batch  # noqa: E402
from sgs.synth import default_medians  # noqa: E402

if __name__ == '__main__':
    U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    _lib.ensure_init(0)
    med = torch.from_numpy(default_medians(40, 9)).cuda()
    g = torch.Generator(device='cuda'); g.manual_seed(3000)
    idx = torch.randint(0, 9, (U, T, 40), device='cuda', generator=g)
    spec = torch.gather(med[None, None].expand(U, T, 40, 9), 3, idx[..., None])[..., 0].contiguous()   # per-bin logistic medians
    noise = torch.rand((U, 160 * (T - 1) + 800), dtype=torch.float64, device='cuda', generator=g)
    for _ in range(2):
        pcm = griffin_lim_batch(spec, noise, num_iterations=iters)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        pcm = griffin_lim_batch(spec, noise, num_iterations=iters)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    # 5 N log2 N per complex transform of N/2 = 400 points, one forward + one inverse real transform per frame-iteration
    flop = U * T * iters * 2 * 2.5 * 800 * np.log2(800)
    print(json.dumps({"utterances": U, "frames": T, "iterations": iters, "ms": ms, "audio_seconds_per_s": U * T * 0.01 / (ms * 1e-3),
                      "frame_iterations_per_s": U * T * iters / (ms * 1e-3), "fft_gflops_nominal": flop / (ms * 1e-3) / 1e9,
                      "checksum": int(pcm.to(torch.int64).abs().sum().item())}))
