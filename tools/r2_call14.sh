#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call14
mkdir -p "$out"; : > "$out/summary.txt"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "reconstruct" > "$out/recon_tests.log" 2>&1; echo "recon tests rc=$?" | tee -a "$out/summary.txt"
tail -n 15 "$out/recon_tests.log"
timeout 300 python tools/recon_bench.py > "$out/recon_bench.jsonl" 2> "$out/recon_bench.err"; echo "recon bench rc=$?" | tee -a "$out/summary.txt"
cat "$out/recon_bench.jsonl"; tail -n 3 "$out/recon_bench.err"
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_linear_host.py -q -m gpu > "$out/fullsize.log" 2>&1; echo "fullsize rc=$?" | tee -a "$out/summary.txt"
tail -n 4 "$out/fullsize.log"
