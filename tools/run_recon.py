"""Run reconstruct_had on one shape a few times (for ncu captures).  python tools/run_recon.py k n K [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exllamav3_b200 import ext
k, n, K = [int(v) for v in sys.argv[1:4]]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
Bs = [torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(iters)]
su = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); sv = torch.randn(n, generator=g, device=dev).half()
w = torch.empty((k, n), dtype=torch.half, device=dev)
for i in range(iters):
    ext.reconstruct_had_slice(w, Bs[i], su, sv, K, False, True, 0)
torch.cuda.synchronize()
print("done", float(w.float().abs().mean()))
