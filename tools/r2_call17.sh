#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call17
mkdir -p "$out"; : > "$out/summary.txt"
timeout 600 python -m pytest tests/test_fanout.py tests/test_gpu_fullsize.py -q -m gpu -x > "$out/fanout_tests.log" 2>&1; echo "fanout tests rc=$?" | tee -a "$out/summary.txt"
tail -n 25 "$out/fanout_tests.log"
EXL3B_BENCH_NO_REF_CUDA=1 timeout 900 python bench.py --no-cpu-baseline --steps 200 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
tail -n 5 "$out/bench_n1.err"
python -c "
import json
d=[json.loads(l) for l in open('$out/bench_n1.json') if l.startswith('{')][-1]
print(round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['frac'], d['e2e'])
print(json.dumps(d['reference_launch_list']))
print(d['gpu_launches'], d['config']['workload'][:200])
"
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
tail -n 4 "$out/gpu_suite.log"
