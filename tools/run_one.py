"""Run one exl3_gemm shape a few times (for ncu captures).  python tools/run_one.py k n K m [iters] [path]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exllamav3_b200 import ext
k, n, K, m = [int(v) for v in sys.argv[1:5]]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 4
if len(sys.argv) > 6:
    ext.set_gemm_path({"tc": ext.EXL3B_TAG_TC, "simt": ext.EXL3B_TAG_SIMT, "i8": ext.EXL3B_TAG_TC_I8, "chain": ext.EXL3B_TAG_TC_I8_CHAIN}[sys.argv[6]])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
Bs = [torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(iters)]
su = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); sv = torch.randn(n, generator=g, device=dev).half()
A = torch.randn((m, k), generator=g, device=dev).half(); Ah = torch.empty_like(A)
C = torch.empty((m, n), dtype=torch.half, device=dev)
for i in range(iters):
    ext.exl3_gemm(A, Bs[i], C, su, Ah, sv, -1, False, True, 0)
torch.cuda.synchronize()
print("done", float(C.float().abs().mean()))
