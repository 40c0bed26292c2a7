"""Scratch benchmark of the feature kernels (device-resident input, CUDA-event timing)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
import torch
from sgs import _lib
from sgs.features import FeatureExtractor

ap = argparse.ArgumentParser()
ap.add_argument('--sessions', type=int, default=8)
ap.add_argument('--channels', type=int, default=128)
ap.add_argument('--sr', type=int, default=2048)
ap.add_argument('--dur', type=float, default=600.0)
ap.add_argument('--chunks', type=str, default='auto')
ap.add_argument('--reps', type=int, default=3)
a = ap.parse_args()
_lib.ensure_init(0)
T = int(a.dur * a.sr)
x = torch.empty((a.sessions, T, a.channels), dtype=torch.float32, device='cuda')
for s in range(a.sessions):
    x[s].normal_(0, 50.0)
fe = FeatureExtractor(a.sr)
for ch in a.chunks.split(','):
    chunks = None if ch == 'auto' else int(ch)
    plan = fe.scan_plan(T, a.sessions * a.channels, chunks)
    fe.log_power(x, chunks=chunks); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for r in range(a.reps):
        ev0.record(); out = fe.log_power(x, chunks=chunks); ev1.record(); torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    samples = a.sessions * T * a.channels
    cost = 1.0 + (plan[2] / plan[1] if plan[0] > 1 else 0.0)
    print('chunks=%s plan=(K=%d L=%d W=%d phi=%s) %.2f ms  %.1f Gsamples/s  in=%.0f GB/s  ch-s/s=%.3g  DFMA-equiv=%.2f T/s (x%.2f work)' % (
        ch, plan[0], plan[1], plan[2], plan[3] is not None, best, samples / best / 1e6, samples * 4 / best / 1e6,
        a.sessions * a.channels * a.dur / (best / 1e3), samples * 100 * cost / best / 1e9, cost), flush=True)
    _lib.profile_enable(True)
    fe.log_power(x, chunks=chunks); torch.cuda.synchronize()
    print('   classes (ms): ' + ', '.join('%s %.3f' % (k, _lib.profile_read(k)[0]) for k in
                                          ('iir_init', 'iir_pieces_tail', 'iir_pieces_state', 'iir_pieces_feat', 'iir_state', 'iir_feat')), flush=True)
    _lib.profile_enable(False)
t = fe.tail()
if t is not None:
    print('tail: near %d far %d modes %d lens %s' % (t.near_len, t.far_len, t.n_modes, sorted(set(t.mode_len.tolist()))))
