"""Run the q + k + v fan-out exl3_mgemm (4096 -> 4096 / 1024 / 1024, K = 4, one row) a few times (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exllamav3_b200 import ext
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
k, K, widths = 4096, 4, [4096, 1024, 1024]
sets = []
for i in range(iters):
    ts = [(torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16),
           (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(), torch.randn(n, generator=g, device=dev).half()) for n in widths]
    sets.append((ts, [torch.tensor([t[j].data_ptr() for t in ts], dtype=torch.long, device=dev) for j in range(3)]))
x = torch.randn((1, 1, k), generator=g, device=dev).half()
outs = [torch.empty((1, n), dtype=torch.half, device=dev) for n in widths]
C = torch.empty((3, 1, max(widths)), dtype=torch.half, device=dev); Ah = torch.empty((3, 1, k), dtype=torch.half, device=dev)
snl = torch.tensor(widths, dtype=torch.int, device=dev); cp = torch.tensor([o.data_ptr() for o in outs], dtype=torch.long, device=dev)
for ts, ptrs in sets:
    tag = ext.exl3_mgemm(x, ptrs[0], C, ptrs[1], Ah, ptrs[2], None, None, K, -1, False, True, -1, -1, 0, 1, snl, cp)
torch.cuda.synchronize()
print("done", tag, float(outs[0].float().abs().mean()))
