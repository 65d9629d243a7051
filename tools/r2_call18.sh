#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call18
mkdir -p "$out"; : > "$out/summary.txt"
ids=$(python -m pytest tests/test_fanout.py -m gpu --collect-only -q 2>/dev/null | grep "::")
for id in $ids; do
  timeout 300 python -m pytest "$id" -q -m gpu > "$out/t.log" 2>&1; rc=$?
  echo "$id rc=$rc" | tee -a "$out/summary.txt"
  if [ $rc -ne 0 ]; then grep -E "^E  |Error|assert" "$out/t.log" | head -n 6 | tee -a "$out/summary.txt"; fi
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_fanout.py::test_fanout_equals_separate_calls_and_oracle[1-8192-widths1-4]" -q -m gpu -x > "$out/sanitizer.log" 2>&1
echo "sanitizer rc=$?" | tee -a "$out/summary.txt"
grep -A25 "Invalid\|Error" "$out/sanitizer.log" | head -n 80
