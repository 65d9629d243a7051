#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call8
mkdir -p "$out"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k hgemm -x > "$out/hgemm_tests.log" 2>&1; echo "hgemm tests rc=$?" | tee "$out/summary.txt"
tail -n 6 "$out/hgemm_tests.log"
timeout 300 python tools/hgemm_debug.py > "$out/hgemm_cases.log" 2>&1; echo "hgemm cases rc=$?" | tee -a "$out/summary.txt"
tail -n 20 "$out/hgemm_cases.log" | cut -c1-300
timeout 300 python tools/hgemm_debug.py --time > "$out/hgemm_time.jsonl" 2>&1; echo "hgemm time rc=$?" | tee -a "$out/summary.txt"
cat "$out/hgemm_time.jsonl"
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "suite rc=$?" | tee -a "$out/summary.txt"
tail -n 6 "$out/gpu_suite.log"
timeout 300 python tools/prefill_bench.py > "$out/prefill.jsonl" 2>&1; echo "prefill rc=$?" | tee -a "$out/summary.txt"
cat "$out/prefill.jsonl" | cut -c1-200
