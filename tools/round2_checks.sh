#!/usr/bin/env bash
# First GPU call of round 2 (run under gpurun from the repo root; everything lands in gpurun_out/r02_first/):
#   gpurun --timeout 1500 -- 'bash tools/round2_checks.sh'
# 1. the verified suite (must stay green), 2. the kernels written blind at the end of round 1 (EXL3B_TEST_UNVERIFIED=1,
# under a timeout, never -x: every failure is wanted), 3. the bench line, 4. smoke.
# Two-GPU follow-up (separate call):  gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_checks.sh tp'
set -u
out=gpurun_out/r02_first
mkdir -p "$out"
if [ "${1:-}" = "tp" ]; then
    EXL3B_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_tp_fused.py -q -m gpu -k two_gpus > "$out/tp_two_gpus.log" 2>&1
    echo "two_gpus rc=$?" | tee -a "$out/summary.txt"
    for fl in "" "--fused-allreduce"; do
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
            bench.py --gpus 2 --steps 20 --warmup 3 $fl > "$out/bench_n2${fl:+_fused}.json" 2> "$out/bench_n2${fl:+_fused}.err"
        echo "bench n2 $fl rc=$?" | tee -a "$out/summary.txt"
    done
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
        tools/tp_rowparallel_bench.py > "$out/tp_rowparallel.jsonl" 2> "$out/tp_rowparallel.err"
    echo "row-parallel NCCL vs fused rc=$?" | tee -a "$out/summary.txt"
    cat "$out/tp_rowparallel.jsonl"
    exit 0
fi
timeout 900 python -m pytest tests -q -m gpu -x > "$out/gpu_suite.log" 2>&1;  echo "gpu suite rc=$?" | tee "$out/summary.txt"
EXL3B_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_tp_fused.py -q -m gpu > "$out/unverified.log" 2>&1
echo "unverified (fused all-reduce: world 1, loop-back) rc=$?" | tee -a "$out/summary.txt"
EXL3B_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_moe_routed.py -q -m gpu > "$out/unverified_moe.log" 2>&1
echo "unverified (routed int8 mgemm) rc=$?" | tee -a "$out/summary.txt"
timeout 600 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err";         echo "bench rc=$?" | tee -a "$out/summary.txt"
timeout 300 python __graft_entry__.py --smoke > "$out/smoke.log" 2>&1;             echo "smoke rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/gpu_suite.log" "$out/unverified.log" "$out/unverified_moe.log"
cat "$out/bench_n1.json"
