#!/usr/bin/env bash
# prologue load batching A/B (main = 8 tasks per step, _pb2 = 2), then parity of the kept build
set -u
out=gpurun_out/r02_call23
mkdir -p "$out"; : > "$out/summary.txt"
for v in "" _pb2 "" _pb2; do
  EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200$v.so EXL3B_BENCH_NO_REF_CUDA=1 timeout 600 python bench.py --no-cpu-baseline --steps 200 > "$out/bench$v.json" 2> "$out/bench$v.err"; echo "bench '$v' rc=$?" | tee -a "$out/summary.txt"
  python - "$out/bench$v.json" <<'PY'
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
print(sys.argv[1], round(d['value'], 1), 'tok/s', round(d['ms_per_step'], 4), 'ms; ref list', round(d['reference_launch_list']['value'], 1), '| us:',
      {k: v['us_per_launch'] for k, v in d['qgemm']['decode_hbm'].items()}, d['qgemm'].get('decode_codebooks'))
PY
done
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
tail -n 4 "$out/gpu_suite.log"
