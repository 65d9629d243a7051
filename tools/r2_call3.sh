#!/usr/bin/env bash
# Round 2, GPU call 3: first run of the chain kernel (parity tests under a timeout, then timings)
set -u
out=gpurun_out/r02_call3
mkdir -p "$out"
timeout 600 python -m pytest tests/test_chain.py -q -m gpu -x > "$out/chain_tests.log" 2>&1; echo "chain tests rc=$?" | tee "$out/summary.txt"
tail -n 30 "$out/chain_tests.log"
timeout 900 python tools/chain_bench.py > "$out/chain_bench.log" 2>&1; echo "chain bench rc=$?" | tee -a "$out/summary.txt"
tail -n 25 "$out/chain_bench.log"
