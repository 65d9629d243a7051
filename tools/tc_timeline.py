"""Per-CTA timeline of the tcgen05 kernel (bring-up aid).  python tools/tc_timeline.py k n K m"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from exllamav3_b200 import ext
k, n, K, m = [int(v) for v in sys.argv[1:5]]
knob = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
iters = 6
Bs = [torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(iters)]
su = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); sv = torch.randn(n, generator=g, device=dev).half()
A = torch.randn((m, k), generator=g, device=dev).half(); Ah = torch.empty_like(A)
C = torch.empty((m, n), dtype=torch.half, device=dev)
ext.set_gemm_path(ext.EXL3B_TAG_TC)
ext.lib.exl3b_debug_tc_knob(knob)
dbg = torch.zeros((148, 16), dtype=torch.int64, device=dev)
ext.lib.exl3b_debug_tc_timeline.argtypes = [ctypes.c_void_p]
for i in range(iters):
    if i == iters - 1:
        ext.lib.exl3b_debug_tc_timeline(dbg.data_ptr())
    ext.exl3_gemm(A, Bs[i], C, su, Ah, sv, -1, False, True, 0)
torch.cuda.synchronize()
ext.lib.exl3b_debug_tc_timeline(None)
d = dbg.cpu().numpy().astype(np.int64)
d = d[d[:, 0] > 0]
t0 = d[:, 0].min()
names = {0: "entry", 1: "setup done", 2: "producer first batch issued", 3: "decode: first W_FULL", 4: "decode: first unit done",
         5: "mma: first operands ready", 6: "mma: last commit", 10: "decode: last unit done", 7: "epi: first D_FULL",
         8: "epi: done", 9: "teardown"}
print(f"knob={knob} shape k={k} n={n} K={K} m={m}: {len(d)} CTAs; ns relative to the earliest CTA entry (min / median / max)")
for s in (0, 1, 2, 3, 4, 5, 10, 6, 7, 8, 9):
    v = d[:, s] - t0
    v = v[d[:, s] > 0]
    if len(v):
        print(f"  {names[s]:32s} {v.min():8d} {int(np.median(v)):8d} {v.max():8d}")
