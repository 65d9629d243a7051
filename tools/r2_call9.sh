#!/usr/bin/env bash
# Round 2, GPU call 9 (gpurun --gpus 8): tensor-parallel scaling with the fused row-parallel sum as default
set -u
out=gpurun_out/r02_call9
mkdir -p "$out"
run() {  # name nproc args...
  name=$1; np=$2; shift 2
  if [ "$np" = "1" ]; then
    timeout 400 python bench.py --gpus 1 "$@" > "$out/$name.json" 2> "$out/$name.err"
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29$((600 + np)) \
        bench.py --gpus $np "$@" > "$out/$name.json" 2> "$out/$name.err"
  fi
  echo "$name rc=$?" | tee -a "$out/summary.txt"
  python - "$out/$name.json" <<'PY'
import json, sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
    print('   ', d['n_gpus'], 'gpus', round(d['value'],1), 'tok/s', round(d['ms_per_step'],3), 'ms', d['config']['row_parallel_sum'][:30], 'frac', round(d['roofline']['frac'],3))
except Exception as e:
    print('    no line:', e)
PY
}
: > "$out/summary.txt"
run b8_n8 8 --no-cpu-baseline --no-qgemm
run b8_n4 4 --no-cpu-baseline --no-qgemm
run b8_n8_nccl 8 --no-cpu-baseline --no-qgemm --nccl-allreduce --steps 100
run b70_n8 8 --model llama-3.1-70b --steps 100 --no-cpu-baseline --no-qgemm
run b70_n4 4 --model llama-3.1-70b --steps 100 --no-cpu-baseline --no-qgemm
run b70_n1 1 --model llama-3.1-70b --steps 50 --no-cpu-baseline --no-qgemm
tail -n 3 "$out"/*.err | tail -n 30
