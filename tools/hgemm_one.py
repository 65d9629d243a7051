"""Run the dense tcgen05 GEMM on one shape a few times (for ncu tensor-pipe counters).  python tools/hgemm_one.py m k n [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exllamav3_b200 import ext
m, k, n = [int(v) for v in sys.argv[1:4]]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
a = (torch.randn((m, k), device=dev) * 0.5).half(); b = (torch.randn((k, n), device=dev) * 0.05).half()
c = torch.empty((m, n), dtype=torch.half, device=dev)
for _ in range(iters):
    ext.hgemm(a, b, c)
torch.cuda.synchronize()
print("done", float(c.float().abs().mean()))
