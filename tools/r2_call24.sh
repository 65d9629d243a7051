#!/usr/bin/env bash
# final validation of the round: suite, smoke, the default bench line, launch list of one eager token, ncu captures of the two new kernels
set -u
out=gpurun_out/r02_call24
mkdir -p "$out"; : > "$out/summary.txt"
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
tail -n 3 "$out/gpu_suite.log"
timeout 300 python __graft_entry__.py --smoke > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/summary.txt"
tail -n 1 "$out/smoke.log"
timeout 900 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
python -c "
import json
d=[json.loads(l) for l in open('$out/bench_n1.json') if l.startswith('{')][-1]
print(round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['frac'], d['e2e']['value'], d['e2e']['eager_value'])
print(json.dumps(d['reference_launch_list']))
print(json.dumps(d['qgemm'].get('reference_cuda'))[:1500])
print(json.dumps(d['qgemm'].get('decode_codebooks')))
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|had_tiled|mgemm" -c 260 --csv \
    --log-file "$out/launches.csv" python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-qgemm > "$out/launches.log" 2>&1
echo "launch list rc=$?" | tee -a "$out/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reconstruct_had_tc -s 2 -c 1 -o "$out/prof_recon" \
    python tools/run_recon.py 4096 14336 4 4 > "$out/prof_recon.log" 2>&1
echo "ncu recon rc=$?" | tee -a "$out/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_i8 -s 2 -c 1 -o "$out/prof_fanout" \
    python tools/run_fanout.py 4 > "$out/prof_fanout.log" 2>&1
echo "ncu fanout rc=$?" | tee -a "$out/summary.txt"
ls -la "$out"
