#!/usr/bin/env bash
# Round 2, GPU call 7: the round's profile artefacts + full bench line.
set -u
out=gpurun_out/r02_call7
mkdir -p "$out"
timeout 900 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee "$out/summary.txt"
tail -c 3000 "$out/bench_n1.json"; echo
# launch list of two eager tokens (device time per launch, cold-cache / serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|had_tiled|mgemm|chain" -c 330 --csv \
    --log-file "$out/launches.csv" python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-qgemm > "$out/launches.log" 2>&1
echo "launch list rc=$?" | tee -a "$out/summary.txt"
# --set full captures of the dominant kernel on three shapes, and of the chain kernel
for spec in "4096 4096 4 1 q" "4096 14336 4 1 gate" "4096 128256 6 1 lm_head"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_i8 -s 2 -c 1 -o "$out/prof_$5" \
      python tools/run_one.py $1 $2 $3 $4 4 > "$out/prof_$5.log" 2>&1
  echo "ncu full $5 rc=$?" | tee -a "$out/summary.txt"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_i8 -s 2 -c 1 -o "$out/prof_chain_gate" \
    python tools/run_one.py 4096 14336 4 1 4 chain > "$out/prof_chain_gate.log" 2>&1
echo "ncu full chain rc=$?" | tee -a "$out/summary.txt"
# prefill: tensor-pipe counters of the dense tcgen05 GEMM
for spec in "65536 4096 4096" "16384 4096 14336"; do
  set -- $spec
  timeout 600 ncu --metrics sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed \
      --clock-control none -k regex:hgemm_tc -s 1 -c 1 --csv --log-file "$out/hgemm_tensor_$1x$2x$3.csv" python tools/hgemm_one.py $1 $2 $3 > "$out/hgemm_$1.log" 2>&1
  echo "ncu hgemm $1x$2x$3 rc=$?" | tee -a "$out/summary.txt"
done
timeout 600 python tools/sweep_grid.py > "$out/sweep_grid.jsonl" 2>&1; echo "sweep rc=$?" | tee -a "$out/summary.txt"
cat "$out/sweep_grid.jsonl"
# K = 4 branch-free decode experiment (second build of the same library)
EXL3B_LIBRARY=$PWD/exllamav3_b200/libexl3b200_bf.so timeout 600 python bench.py --no-cpu-baseline --no-qgemm > "$out/bench_branchfree.json" 2> "$out/bench_branchfree.err"
echo "bench branchfree rc=$?" | tee -a "$out/summary.txt"
python -c "
import json
for f in ('bench_n1','bench_branchfree'):
    d=[json.loads(l) for l in open('$out/'+f+'.json') if l.startswith('{')][-1]
    print(f, round(d['value'],1), round(d['ms_per_step'],4), d['e2e'].get('value'), d['e2e'].get('eager_value'), d['clocks'])
"
ls -la "$out"
