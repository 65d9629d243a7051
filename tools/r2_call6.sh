#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call6
mkdir -p "$out"
timeout 600 python -m pytest tests/test_chain.py -q -m gpu -x > "$out/chain_tests.log" 2>&1; echo "chain tests rc=$?" | tee "$out/summary.txt"
tail -n 4 "$out/chain_tests.log"
for shp in "4096 14336 4" "4096 4096 4"; do
  EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200_dbg.so timeout 300 python tools/chain_timeline.py $shp > "$out/timeline_$(echo $shp | tr ' ' '_').log" 2>&1
  echo "timeline $shp rc=$?" | tee -a "$out/summary.txt"
  cat "$out/timeline_$(echo $shp | tr ' ' '_').log"
done
timeout 900 python tools/chain_bench.py > "$out/chain_bench.log" 2>&1; echo "chain bench rc=$?" | tee -a "$out/summary.txt"
tail -n 14 "$out/chain_bench.log" | cut -c1-300
for md in blocks layer; do
  timeout 600 python bench.py --mode $md --no-cpu-baseline --no-qgemm > "$out/bench_$md.json" 2> "$out/bench_$md.err"; echo "bench $md rc=$?" | tee -a "$out/summary.txt"
  python -c "
import json
d=[json.loads(l) for l in open('$out/bench_$md.json') if l.startswith('{')][-1]
print('$md', d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"
done
