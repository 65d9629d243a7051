#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call19
mkdir -p "$out"; : > "$out/summary.txt"
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
tail -n 6 "$out/gpu_suite.log"
timeout 300 python __graft_entry__.py --smoke > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/summary.txt"
tail -n 3 "$out/smoke.log"
timeout 900 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
tail -n 5 "$out/bench_n1.err"
python -c "
import json
d=[json.loads(l) for l in open('$out/bench_n1.json') if l.startswith('{')][-1]
print(round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['frac'], d['e2e'])
print(json.dumps(d['reference_launch_list']))
print(json.dumps(d['cpu_baseline']))
print(json.dumps(d['qgemm'].get('reference_cuda'))[:600])
"
