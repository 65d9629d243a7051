#!/usr/bin/env bash
# fan-out launches with device-patched tensor maps (main) against row copies (_rc3)
set -u
out=gpurun_out/r02_call21
mkdir -p "$out"; : > "$out/summary.txt"
timeout 600 python -m pytest tests/test_fanout.py -q -m gpu > "$out/fanout_tests.log" 2>&1; echo "fanout tests rc=$?" | tee -a "$out/summary.txt"
tail -n 8 "$out/fanout_tests.log"
for v in "" _rc3 ""; do
  EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200$v.so EXL3B_BENCH_NO_REF_CUDA=1 timeout 600 python bench.py --no-cpu-baseline --steps 200 > "$out/bench$v.json" 2> "$out/bench$v.err"; echo "bench '$v' rc=$?" | tee -a "$out/summary.txt"
  python - "$out/bench$v.json" <<'PY'
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
print(sys.argv[1], round(d['value'], 1), 'tok/s', round(d['ms_per_step'], 4), 'ms; ref list', round(d['reference_launch_list']['value'], 1))
PY
done
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
tail -n 4 "$out/gpu_suite.log"
