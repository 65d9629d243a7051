"""Prefill-side timing of the EXL3 linear (rows > 144: reconstruct[_had] + dense GEMM, modules/quant/exl3.py:161-218)
and of the fused decode-GEMM kernel at mid m.  python tools/prefill_bench.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exllamav3_b200 import ext, QLinear
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
peak = 1673.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass
for (k, n, K) in ((4096, 4096, 4), (4096, 14336, 4), (14336, 4096, 4)):
    tr = torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16)
    suh = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); svh = torch.randn(n, generator=g, device=dev).half()
    lin = QLinear(tr, suh, svh, mul1=True)
    for m in (16, 64, 128, 256, 1024, 2048, 8192, 65536):
        if m * max(k, n) * 2 > 6e9:
            continue
        x = torch.randn((m, k), generator=g, device=dev).half()
        for _ in range(2): y = lin.forward(x, {})
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 5
        e0.record()
        for _ in range(it): y = lin.forward(x, {})
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / it
        tf = 2.0 * m * k * n / ms / 1e9
        path = "kernel" if m <= 144 else ("reconstruct_had+hgemm" if m >= 1024 else "had+reconstruct+hgemm+had")
        print(json.dumps(dict(k=k, n=n, K=K, m=m, path=path, ms=round(ms, 4), tflops=round(tf, 1), frac_of_measured_bf16_peak=round(tf / peak, 3))), flush=True)
        del x, y
