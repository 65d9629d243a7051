#!/usr/bin/env bash
# two GPUs: the multigpu test, and the 8B / 70B tokens with the fan-out launch list (default) and without it
set -u
out=gpurun_out/r02_call22
mkdir -p "$out"; : > "$out/summary.txt"
timeout 600 python -m pytest tests -q -m multigpu > "$out/multigpu.log" 2>&1; echo "multigpu tests rc=$?" | tee -a "$out/summary.txt"
tail -n 3 "$out/multigpu.log"
run() {  # name nproc args...
  name=$1; np=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29$((600 + np)) \
      bench.py --gpus $np "$@" > "$out/$name.json" 2> "$out/$name.err"
  echo "$name rc=$?" | tee -a "$out/summary.txt"
  python - "$out/$name.json" <<'PY'
import json, sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
    print('   ', d['n_gpus'], 'gpus', round(d['value'],1), 'tok/s', round(d['ms_per_step'],3), 'ms', 'ref list', d.get('reference_launch_list') and round(d['reference_launch_list']['value'],1), 'frac', round(d['roofline']['frac'],3))
except Exception as e:
    print('    no line:', e)
PY
}
run b8_n2 2 --no-cpu-baseline --no-qgemm --steps 200
run b8_n2_nofan 2 --no-cpu-baseline --no-qgemm --steps 200 --no-fanout
run b70_n2 2 --model llama-3.1-70b --steps 100 --no-cpu-baseline --no-qgemm
tail -n 3 "$out"/*.err | tail -n 20
