#!/usr/bin/env bash
# weights as row copies instead of 2-D tensor-map boxes: 0 = fan-out launches only (main), _rc1 = all multi-matrix launches, _rc2 = all launches
set -u
out=gpurun_out/r02_call20
mkdir -p "$out"; : > "$out/summary.txt"
for v in "" _rc1 _rc2 ""; do
  EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200$v.so EXL3B_BENCH_NO_REF_CUDA=1 timeout 600 python bench.py --no-cpu-baseline --steps 200 > "$out/bench$v.json" 2> "$out/bench$v.err"; echo "bench '$v' rc=$?" | tee -a "$out/summary.txt"
  python - "$out/bench$v.json" <<'PY'
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
print(sys.argv[1], round(d['value'], 1), 'tok/s', round(d['ms_per_step'], 4), 'ms; ref list', round(d['reference_launch_list']['value'], 1), '| us:',
      {k: v['us_per_launch'] for k, v in d['qgemm']['decode_hbm'].items()})
PY
done
EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200_rc2.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fanout.py tests/test_gpu_fullsize.py tests/test_tp_fused.py tests/test_moe_routed.py -q -m gpu > "$out/tests_rc2.log" 2>&1; echo "tests rc2 rc=$?" | tee -a "$out/summary.txt"
tail -n 3 "$out/tests_rc2.log"
