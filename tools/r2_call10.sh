#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call10
mkdir -p "$out"; : > "$out/summary.txt"
for v in "" _e2 _e3 _e23; do
  lib=$PWD/exllamav3_b200/libexl3b200$v.so
  EXL3B_LIBRARY=$lib timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_tp_fused.py tests/test_moe_routed.py -q -m gpu -x -k "i8 or fused or routed or mgemm" > "$out/tests$v.log" 2>&1; echo "tests[$v] rc=$?" | tee -a "$out/summary.txt"
  tail -n 2 "$out/tests$v.log"
  EXL3B_LIBRARY=$lib timeout 300 python bench.py --steps 200 --no-cpu-baseline --no-qgemm > "$out/bench$v.json" 2> "$out/bench$v.err"; echo "bench[$v] rc=$?" | tee -a "$out/summary.txt"
  python -c "
import json
d=[json.loads(l) for l in open('$out/bench$v.json') if l.startswith('{')][-1]
print('   variant [$v]:', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms')"
done
# 5..8 rows: in-kernel transform vs tiled path on the q shape
for thr in 0 2 8; do
  EXL3B_FUSED_X_UNITS=$thr timeout 300 python - <<'PY' 2>&1 | tail -n 2
import os, sys, torch
sys.path.insert(0, os.getcwd())
from exllamav3_b200 import ext
dev = torch.device("cuda:0"); g = torch.Generator(device=dev); g.manual_seed(0)
for (k, n) in ((4096, 4096), (4096, 14336)):
    for m in (5, 8):
        trs = [torch.randint(0, 65536, (k // 16, n // 16, 64), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(16)]
        su = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); sv = torch.randn(n, generator=g, device=dev).half()
        x = torch.randn((m, k), generator=g, device=dev).half(); xh = torch.empty_like(x); y = torch.empty((m, n), dtype=torch.half, device=dev)
        def run():
            for t in trs: ext.exl3_gemm(x, t, y, su, xh, sv, -1, False, True, 0)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st): run(); run()
        st.synchronize(); gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st); [gr.replay() for _ in range(5)]; e1.record(st)
        e1.synchronize()
        print(f"thr={os.environ['EXL3B_FUSED_X_UNITS']} {k}x{n} m={m}: {e0.elapsed_time(e1) * 1e3 / 80:.2f} us", end=" | ")
print()
PY
done
