#!/usr/bin/env bash
set -u
out=gpurun_out/r02_call11
mkdir -p "$out"; : > "$out/summary.txt"
for thr in 0 8 1000; do
  EXL3B_FUSED_X_UNITS=$thr timeout 300 python - <<'PY' 2>&1 | tail -n 2 | tee -a "$out/m58.log"
import os, sys, torch
sys.path.insert(0, os.getcwd())
from exllamav3_b200 import ext
dev = torch.device("cuda:0"); g = torch.Generator(device=dev); g.manual_seed(0)
for (k, n) in ((4096, 4096), (4096, 14336), (14336, 4096)):
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
for thr in 0 1000; do
  EXL3B_FUSED_X_UNITS=$thr timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -m gpu -k "batch_8 or vs_oracle or ragged_rows" > "$out/tests_thr$thr.log" 2>&1; echo "tests thr=$thr rc=$?" | tee -a "$out/summary.txt"
  tail -n 2 "$out/tests_thr$thr.log"
done
