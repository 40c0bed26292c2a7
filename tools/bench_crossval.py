"""Many-fold re-train / decode (eval_steps/exp1.py:54-160: 10 folds, x 100 randomised runs for the
chance level) on the device against the CPU port.

    python tools/bench_crossval.py [channels] [sample_rate] [seconds] [folds]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29556 tools/bench_crossval.py ...

GPU: local.crossval.cross_validate (recording uploaded once, folds dealt over the ranks).  CPU: the oracle's train +
streaming decode of ONE fold on one core, scaled by the number of folds (folds are equal in size)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

if __name__ == '__main__':
    import torch
    import torch.distributed as dist
    from sgs import _lib, synth
    from local import crossval
    n_ch, sr, seconds, folds = (int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 1024,
                                float(sys.argv[3]) if len(sys.argv) > 3 else 300.0, int(sys.argv[4]) if len(sys.argv) > 4 else 10)
    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    _lib.ensure_init(local)
    eeg = synth.seeg_session(5, n_ch, sr, seconds)
    audio = synth.audio_session(5, seconds)
    crossval.cross_validate(eeg[:int(30 * sr)], audio[:30 * 16000], sr, 16000, [], nb_folds=2)        # plans, pools
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reco, orig, wav, (mean, std, rs) = crossval.cross_validate(eeg, audio, sr, 16000, [], norm_factor=10, nb_folds=folds)
    gpu_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([gpu_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gpu_s = float(t.item())
    out = {"workload": "%d-fold re-train + decode of a %d ch x %g s @ %d Hz session (eval_steps/exp1.py)" % (folds, n_ch, seconds, sr),
           "gpus": world, "gpu_wall_s": gpu_s, "stages_rank0": {k: round(v, 4) if isinstance(v, float) else v for k, v in crossval.last_profile.items()},
           "pearson_r_mean": float(mean), "frames": int(len(reco))}
    if rank == 0:
        import oracle as O
        (k, (e0, e1), (a0, a1), e_end, a_end) = crossval.fold_bounds(len(eeg), len(audio), sr, 16000, folds)[folds // 2]
        t0 = time.perf_counter()
        x_train = np.vstack([eeg[:e0], eeg[e1:e_end]]).astype(np.float64)
        y_train = np.concatenate([audio[:a0], audio[a1:a_end]])
        _, _, medians, est, select = O.train(x_train, y_train, sr, [])
        x_test = eeg[e0:e1].astype(np.float64)
        n_frames = int((e1 - e0) / sr * 100) + 8                       # at least as many rows as the stretch has 10 ms frames
        noise = np.random.RandomState(k).rand(n_frames, 480)
        O.decode_streaming(x_test, sr, est, select, medians, noise, gl_norm=10, chunk_size=32)
        cpu_fold = time.perf_counter() - t0
        out["cpu_baseline"] = {"seconds_one_fold": cpu_fold, "scaled_to_all_folds_s": cpu_fold * folds, "cores": 1, "kind": "port",
                               "sample": "oracle.train + oracle.decode_streaming of fold %d" % k}
        out["speedup_vs_one_core"] = cpu_fold * folds / gpu_s
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
