#!/usr/bin/env bash
# Round 2, GPU call 12: exact kernel at <= 4 rows (other codebooks): in-kernel transform vs tiled path; then the full validation
set -u
out=gpurun_out/r02_call12
mkdir -p "$out"; : > "$out/summary.txt"
for rows in 4 0; do
  EXL3B_FUSED_X_ROWS=$rows timeout 300 python - <<'PY' 2>&1 | tail -n 2 | tee -a "$out/exact_small_m.log"
import os, sys, torch
sys.path.insert(0, os.getcwd())
from exllamav3_b200 import ext
dev = torch.device("cuda:0"); g = torch.Generator(device=dev); g.manual_seed(0)
for (k, n) in ((4096, 4096), (4096, 14336)):
    for m in (1, 4):
        trs = [torch.randint(0, 65536, (k // 16, n // 16, 64), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(16)]
        su = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); sv = torch.randn(n, generator=g, device=dev).half()
        x = torch.randn((m, k), generator=g, device=dev).half(); xh = torch.empty_like(x); y = torch.empty((m, n), dtype=torch.half, device=dev)
        def run():
            for t in trs: ext.exl3_gemm(x, t, y, su, xh, sv, -1, False, False, 0)      # 3INST codebook -> exact tcgen05 kernel
        st = torch.cuda.Stream()
        with torch.cuda.stream(st): run(); run()
        st.synchronize(); gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st); [gr.replay() for _ in range(5)]; e1.record(st)
        e1.synchronize()
        print(f"fused_rows={os.environ['EXL3B_FUSED_X_ROWS']} cb0 {k}x{n} m={m}: {e0.elapsed_time(e1) * 1e3 / 80:.2f} us", end=" | ")
print()
PY
done
timeout 900 python -m pytest tests -q -m gpu > "$out/gpu_suite.log" 2>&1; echo "gpu suite rc=$?" | tee -a "$out/summary.txt"
tail -n 4 "$out/gpu_suite.log"
timeout 300 python __graft_entry__.py --smoke > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/summary.txt"; tail -n 2 "$out/smoke.log"
timeout 900 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
python -c "
import json
d=[json.loads(l) for l in open('$out/bench_n1.json') if l.startswith('{')][-1]
print(round(d['value'],1), round(d['ms_per_step'],4), d['e2e'], d['clocks'])
print(json.dumps(d['qgemm']['decode_batch']))
print(json.dumps(d['qgemm']['prefill_tensor']))
"
timeout 300 python bench.py --impl reference --steps 3 > "$out/bench_ref.json" 2> "$out/bench_ref.err"; echo "bench reference rc=$?" | tee -a "$out/summary.txt"; tail -c 600 "$out/bench_ref.json"
