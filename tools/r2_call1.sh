#!/usr/bin/env bash
# Round 2, GPU call 1: verified suite, the kernels written blind at the end of round 1 (EXL3B_TEST_UNVERIFIED=1), bench line, smoke,
# and in-kernel timelines of the int8 decode-GEMM (bring-up build libexl3b200_dbg.so) on the three Llama shapes.
set -u
out=gpurun_out/r02_call1
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/gpu.txt" 2>&1
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1;  echo "gpu suite rc=$?" | tee "$out/summary.txt"
EXL3B_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_tp_fused.py -q -m gpu > "$out/unverified_tp.log" 2>&1
echo "unverified (fused all-reduce: world 1, loop-back) rc=$?" | tee -a "$out/summary.txt"
EXL3B_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_moe_routed.py -q -m gpu > "$out/unverified_moe.log" 2>&1
echo "unverified (routed int8 mgemm) rc=$?" | tee -a "$out/summary.txt"
timeout 600 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err";         echo "bench rc=$?" | tee -a "$out/summary.txt"
timeout 300 python __graft_entry__.py --smoke > "$out/smoke.log" 2>&1;             echo "smoke rc=$?" | tee -a "$out/summary.txt"
for shp in "4096 4096 4 1" "4096 14336 4 1" "14336 4096 4 1" "4096 1024 4 1"; do
    nm=$(echo $shp | tr ' ' '_')
    EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200_dbg.so EXL3B_PATH=210 timeout 300 python tools/tc_timeline.py $shp > "$out/timeline_$nm.log" 2>&1
    echo "timeline $shp rc=$?" | tee -a "$out/summary.txt"
done
tail -5 "$out/gpu_suite.log" "$out/unverified_tp.log" "$out/unverified_moe.log"
cat "$out/bench_n1.json"
