"""Where the tail of the streaming latency comes from: the latency leg of bench.py under different hand-over settings of the
Receivers and interpreter switch intervals.  Usage: python tools/latency_tail.py [seconds]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    ("default: time-sharing policy, no hand-over while frames stream", {}, None),
    ("pinned + SCHED_FIFO (decode.realtime)", {"SGS_LAT_RT": "1"}, None),
    ("pinned + SCHED_FIFO, hand-over every 0.25 s on a flusher thread, at most 32 frames", {"SGS_LAT_RT": "1", "SGS_RECEIVER_FLUSH_INTERVAL": "0.25"}, None),
    ("pinned + SCHED_FIFO, round-1 hand-over (0.25 s, any size)", {"SGS_LAT_RT": "1", "SGS_RECEIVER_FLUSH_INTERVAL": "0.25", "SGS_RECEIVER_MAX_BATCH": "0"}, None),
    ("time-sharing, round-1 hand-over", {"SGS_RECEIVER_FLUSH_INTERVAL": "0.25", "SGS_RECEIVER_MAX_BATCH": "0"}, None),
]

if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--child':
        sys.path.insert(0, ROOT)
        import bench
        sw = float(sys.argv[3])
        if sw > 0:
            sys.setswitchinterval(sw)
        r = bench.latency_leg(float(sys.argv[2]), float(sys.argv[2]))
        keep = ("frames", "p50_ms", "p99_ms", "p99.9_ms", "max_ms", "frames_over_1ms")
        res = {k: {q: v[q] for q in keep} for k, v in r.items() if isinstance(v, dict) and "p99_ms" in v}
        res['scheduling'] = r.get('scheduling')
        print(json.dumps(res))
        sys.exit(0)
    seconds = sys.argv[1] if len(sys.argv) > 1 else "30"
    only = os.environ.get("SGS_TAIL_CASES")
    for ci, (name, env, sw) in enumerate(CASES):
        if only and str(ci) not in only.split(","):
            continue
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', seconds, str(sw or 0)], env=e, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith('{')]
        print(json.dumps({"case": name, "result": json.loads(line[-1]) if line else out.stderr[-400:]}), flush=True)
