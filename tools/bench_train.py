"""BASELINE config 3 (LDA fit statistics on 128 ch x 1 h): device time of sgs_lda_stats, tensor-core (kind::i8 digit GEMMs)
against the fp64 CUDA-core kernels, inputs resident in HBM.  Usage: python tools/bench_train.py [rows] [width]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
from sgs import _lib  # noqa: E402


def run(x, sel, lab, nf, nb, reps=5):
    n, width = x.shape
    dev = x.device
    xbar = torch.empty(nf, dtype=torch.float64, device=dev)
    G = torch.empty((nf, nf), dtype=torch.float64, device=dev)
    sums = torch.empty((nb, 9, nf), dtype=torch.float64, device=dev)
    cnt = torch.empty((nb, 9), dtype=torch.float64, device=dev)
    sel_h = np.ascontiguousarray(sel, dtype=np.int32)

    def call():
        _lib.check(_lib.lib().sgs_lda_stats(_lib.ptr(x), n, width, _lib.ptr(sel_h), nf, _lib.ptr(lab), nb, 9, None, _lib.ptr(xbar),
                                            _lib.ptr(G), _lib.ptr(sums), _lib.ptr(cnt), None))
    call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        call()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, G.cpu().numpy(), sums.cpu().numpy()


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 359976
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    nf, nb = 150, 40
    _lib.ensure_init(0)
    torch.manual_seed(0)
    x = torch.randn((n, width), dtype=torch.float64, device='cuda') * 0.6 + 9.0
    lab = torch.randint(0, 9, (n, nb), device='cuda').to(torch.float64)
    sel = np.random.default_rng(0).permutation(width)[:nf]
    out = {"rows": n, "width": width, "features": nf, "bins": nb}
    _lib.profile_enable(True)
    os.environ['SGS_TRAIN_TC'] = '1'
    ms_tc, G_tc, S_tc = run(x, sel, lab, nf, nb)
    k_ms, k_n = _lib.profile_read('train_tc')
    os.environ['SGS_TRAIN_TC'] = '0'
    ms_64, G_64, S_64 = run(x, sel, lab, nf, nb)
    out["lda_stats_ms"] = {"tensor_core_i8": ms_tc, "fp64_cuda_cores": ms_64}
    # integer MACs issued by k_tc_stats: (21 pairs x 2 + 6 x 3) accumulators of 128 x 160 per row, 2 ops per MAC
    ops = 2.0 * (21 * 2 + 6 * 3) * 128 * 160 * (-(-n // 128) * 128)
    out["k_tc_stats"] = {"ms": k_ms / k_n, "int8_tops": ops / (k_ms / k_n * 1e-3) / 1e12,
                         "useful_fp64_equiv_gflop": 2.0 * n * nf * (nf + nb * 9) / 1e9}
    out["max_rel_diff_G"] = float(np.abs(G_tc - G_64).max() / np.abs(G_64).max())
    out["max_rel_diff_sums"] = float(np.abs(S_tc - S_64).max() / np.abs(S_64).max())
    print(json.dumps(out))
