#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call4
mkdir -p "$out"
timeout 600 python -m pytest tests/test_chain.py -q -m gpu > "$out/chain_tests.log" 2>&1; echo "chain tests rc=$?" | tee "$out/summary.txt"
tail -n 12 "$out/chain_tests.log"
for shp in "4096 14336 4" "4096 4096 4"; do
  EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200_dbg.so timeout 300 python tools/chain_timeline.py $shp > "$out/timeline_$(echo $shp | tr ' ' '_').log" 2>&1
  echo "timeline $shp rc=$?" | tee -a "$out/summary.txt"
  cat "$out/timeline_$(echo $shp | tr ' ' '_').log"
done
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu > "$out/fullsize.log" 2>&1; echo "fullsize rc=$?" | tee -a "$out/summary.txt"
tail -n 25 "$out/fullsize.log"
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_fullsize.py --deselect tests/test_chain.py > "$out/gpu_suite.log" 2>&1; echo "suite rc=$?" | tee -a "$out/summary.txt"
tail -n 5 "$out/gpu_suite.log"
