#!/usr/bin/env bash
# Round 2, GPU call 2 (gpurun --gpus 2): refactored suite on one GPU, the two-process fused all-reduce test, bench N=2 NCCL vs fused,
# row-parallel site timing both ways.
set -u
out=gpurun_out/r02_call2
mkdir -p "$out"
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > "$out/gpu.txt" 2>&1
timeout 600 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1;  echo "gpu suite rc=$?" | tee "$out/summary.txt"
timeout 600 python -m pytest tests -q -m multigpu > "$out/multigpu.log" 2>&1; echo "multigpu rc=$?" | tee -a "$out/summary.txt"
for fl in "" "--fused-allreduce"; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        bench.py --gpus 2 --steps 20 --warmup 3 $fl > "$out/bench_n2${fl:+_fused}.json" 2> "$out/bench_n2${fl:+_fused}.err"
    echo "bench n2 $fl rc=$?" | tee -a "$out/summary.txt"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
    tools/tp_rowparallel_bench.py > "$out/tp_rowparallel.jsonl" 2> "$out/tp_rowparallel.err"
echo "row-parallel NCCL vs fused rc=$?" | tee -a "$out/summary.txt"
for fl in "" "--fused-allreduce"; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
        bench.py --gpus 2 --steps 10 --warmup 3 --model llama-3.1-70b $fl > "$out/bench70_n2${fl:+_fused}.json" 2> "$out/bench70_n2${fl:+_fused}.err"
    echo "bench 70b n2 $fl rc=$?" | tee -a "$out/summary.txt"
done
tail -4 "$out/gpu_suite.log" "$out/multigpu.log"
cat "$out"/bench*_n2*.json "$out/tp_rowparallel.jsonl"
