#!/usr/bin/env bash
# decode-ahead A/B (main lib = level 2; _da0 / _da1 = levels 0 / 1), parity of the restructured kernel, tensor-core reconstruct_had splits
set -u
out=gpurun_out/r02_call15
mkdir -p "$out"; : > "$out/summary.txt"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_moe_routed.py tests/test_tp_fused.py tests/test_linear_host.py -q -m gpu -x > "$out/tests_main.log" 2>&1; echo "tests main rc=$?" | tee -a "$out/summary.txt"
tail -n 6 "$out/tests_main.log"
timeout 300 python tools/recon_bench.py > "$out/recon_bench.jsonl" 2> "$out/recon_bench.err"; echo "recon bench rc=$?" | tee -a "$out/summary.txt"
cat "$out/recon_bench.jsonl"; tail -n 3 "$out/recon_bench.err"
for v in "" _da0 _da1 ""; do
  EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200$v.so EXL3B_BENCH_NO_REF_CUDA=1 timeout 600 python bench.py --no-cpu-baseline --steps 200 > "$out/bench$v.json" 2> "$out/bench$v.err"; echo "bench '$v' rc=$?" | tee -a "$out/summary.txt"
  python - "$out/bench$v.json" <<'PY'
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
q = d['qgemm']
print(sys.argv[1], round(d['value'], 1), 'tok/s', round(d['ms_per_step'], 4), 'ms; e2e', round(d['e2e']['value'], 1), '| shapes us:',
      {k: v['us_per_launch'] for k, v in q['decode_hbm'].items()})
PY
done
for shp in "4096 4096 4 1" "4096 14336 4 1"; do
    nm=$(echo $shp | tr ' ' '_')
    EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200_dbg.so EXL3B_PATH=210 timeout 300 python tools/tc_timeline.py $shp > "$out/timeline_$nm.log" 2>&1
    echo "timeline $shp rc=$?" | tee -a "$out/summary.txt"
done
head -n 30 "$out/timeline_4096_4096_4_1.log"
