"""In-kernel timeline of the chain kernel (bring-up build).  EXL3B_LIBRARY=.../libexl3b200_dbg.so python tools/chain_timeline.py k n K [ops]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from exllamav3_b200 import ext
k, n, K = [int(v) for v in sys.argv[1:4]]
nops = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
mats = []
for i in range(nops):
    tr = torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16)
    su = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); sv = torch.randn(n, generator=g, device=dev).half()
    x = torch.randn((1, k), generator=g, device=dev).half(); y = torch.empty((1, n), dtype=torch.float, device=dev)
    mats.append((tr, su, sv, x, y))
ext.lib.exl3b_debug_tc_timeline.argtypes = [ctypes.c_void_p]
chs = [ext.GemmChain([dict(x=x, trellis=tr, suh=su, svh=sv, y=y, mul1=True)]) for (tr, su, sv, x, y) in mats]
dbgs = [torch.zeros((148, 64), dtype=torch.int64, device=dev) for _ in chs]
for c in chs: c.run()
torch.cuda.synchronize()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for c, d in zip(chs, dbgs):
            ext.lib.exl3b_debug_tc_timeline(d.data_ptr())
            c.run()
    ext.lib.exl3b_debug_tc_timeline(None)
    gr.replay(); torch.cuda.synchronize(); gr.replay()
torch.cuda.synchronize()
D = [d.cpu().numpy().astype(np.int64) for d in dbgs]
t0 = min(d[d[:, 61] > 0][:, 61].min() for d in D)
med = lambda v: int(np.median(v)) if len(v) else -1
print(f"chain kernel k={k} n={n} K={K}: {nops} single-op launches, graph replay; ns (median over CTAs)")
for i, d in enumerate(D):
    ok = d[:, 61] > 0
    e = d[ok]
    print(f" launch {i}: entry {med(e[:, 61] - t0)}  scale pass start {med(e[:, 59] - t0)}  scale pass end {med(e[:, 60] - t0)}  exit {med(e[:, 62] - t0)} (max {int((e[:, 62] - t0).max())})")
d = D[-1]
ok = (d[:, 0] > 0)
if ok.any():
    base = d[ok, 0]
    lab = ["top", "W_FULL", "LDS+SHFL", "A_EMPTY", "decode+STTM issued", "wait::st", "arrived"]
    print("  decode warp (group lead), unit 8:", {l: med(d[ok, j] - base) for j, l in enumerate(lab)}, " absolute start:", med(base - t0_last) if False else "")
    okm = d[:, 8] > 0
    print("  mma warp, unit 8 (rel. decode top):", {"loop top": med(d[okm & ok, 8] - d[okm & ok, 0]), "A_FULL": med(d[okm & ok, 9] - d[okm & ok, 0]), "issued+committed": med(d[okm & ok, 10] - d[okm & ok, 0])})
    okx = d[:, 56] > 0
    print("  digit warp, unit 8:", {"W_EMPTY": med(d[okx, 57] - d[okx, 56]), "digits+arrive": med(d[okx, 58] - d[okx, 56]), "start vs decode top": med(d[okx & ok, 56] - d[okx & ok, 0])})
    e = d[ok]
    print("  absolute (vs this launch's entry): decode unit 8 top", med(e[:, 0] - e[:, 61]), " scale pass end", med(e[:, 60] - e[:, 61]), " exit", med(e[:, 62] - e[:, 61]))
