#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call16
mkdir -p "$out"; : > "$out/summary.txt"
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
tail -n 4 "$out/gpu_suite.log"
timeout 300 python tools/recon_bench.py > "$out/recon_bench.jsonl" 2> "$out/recon_bench.err"; echo "recon bench rc=$?" | tee -a "$out/summary.txt"
cat "$out/recon_bench.jsonl"; tail -n 3 "$out/recon_bench.err"
EXL3B_BENCH_NO_REF_CUDA=1 timeout 900 python bench.py --no-cpu-baseline --steps 200 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
python -c "
import json
d=[json.loads(l) for l in open('$out/bench_n1.json') if l.startswith('{')][-1]
print(round(d['value'],1), round(d['ms_per_step'],4))
print(json.dumps(d['qgemm']['reconstruct']))
print(json.dumps(d['qgemm']['prefill_tensor']))
"
