#!/usr/bin/env bash
# Per-round profile artefacts in one GPU call (B200_PROFILING.md recipe).  From the repo root:
#   gpurun --timeout 1500 -- 'bash tools/round_profile.sh r02'
# then, back on the CPU box:  bash tools/round_profile.sh r02 extract      (reads gpurun_out/, writes profiles/r02_*)
# 1. bench line (never under a profiler), 2. launch list of two eager tokens, 3. one `ncu --set full` capture of the dominant
# kernel on the gate shape (4096 x 14336, K = 4, m = 1: the capture bench.py's roofline.traffic is scaled from).
set -u
tag=${1:-r02}
out=gpurun_out/$tag
if [ "${2:-}" = "extract" ]; then
    mkdir -p profiles
    cp "$out/bench_n1.json" "profiles/${tag}_bench_n1.json"
    cp "$out/launches.csv" "profiles/${tag}_launches.csv"
    ncu -i "$out/prof_gate.ncu-rep" --page raw --csv > "$out/prof_gate_raw.csv" 2>/dev/null
    python - "$out/prof_gate_raw.csv" "profiles/${tag}_ncu_i8_gate_4096x14336_K4_m1.csv" <<'PY'
import csv, sys, re
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]             # first captured launch
keep = re.compile(r"dram__bytes_(read|write)\.sum$|dram__throughput|gpu__time_duration\.sum|launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic)"
                  r"|sm__cycles_elapsed\.avg$|sm__pipe_tensor|sm__inst_executed_pipe_(alu|fma|fmaheavy|tensor)|smsp__issue_active\.avg\.pct|sm__warps_active\.avg\.pct_of_peak|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$")
with open(sys.argv[2], "w") as f:
    f.write("metric,unit,value\n")
    for n, u, v in zip(names, units, vals):
        if keep.search(n):
            f.write(f"{n},{u},{v.replace(',', '')}\n")
print("wrote", sys.argv[2])
PY
    echo "wrote profiles/${tag}_*; summarise them in profiles/${tag}_ncu_notes.md and point bench.py's NCU_CAPTURE at the new file"
    exit 0
fi
mkdir -p "$out"
timeout 600 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee "$out/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|had_tiled|mgemm" -c 322 --csv \
    --log-file "$out/launches.csv" python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-qgemm > "$out/launches.log" 2>&1
echo "launch list rc=$?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_i8 -s 2 -c 1 -o "$out/prof_gate" \
    python tools/run_one.py 4096 14336 4 1 4 > "$out/prof_gate.log" 2>&1
echo "ncu full rc=$?" | tee -a "$out/summary.txt"
ls -la "$out"
