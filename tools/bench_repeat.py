"""Runs `python bench.py` N times and prints the per-step times of each run (run-to-run stability of the timed region).
Usage: python tools/bench_repeat.py [N] [tag] [bench.py arguments...]  -> gpurun_out/bench_<tag>_<i>.log"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
tag = sys.argv[2] if len(sys.argv) > 2 else 'repeat'
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
for i in range(1, n + 1):
    log = os.path.join(ROOT, 'gpurun_out', 'bench_%s_%d.log' % (tag, i))
    with open(log, 'w') as fh:
        rc = subprocess.call([sys.executable, os.path.join(ROOT, 'bench.py')] + sys.argv[3:], stdout=fh, stderr=subprocess.STDOUT)
    line = open(log).read().strip().splitlines()[-1]
    try:
        d = json.loads(line)
        lat = d.get('latency') or {}
        clk = d.get('clocks') or {}
        print(rc, round(d['value']), d['ms_each_step'], d.get('host_enqueue_ms_each_step'), d.get('host_stage_ms_first_two_steps'), round(d['e2e']['value']), lat.get('p99_ms'), clk.get('sm_mhz'), clk.get('reasons'), flush=True)
    except ValueError:
        print(rc, line[-300:], flush=True)
