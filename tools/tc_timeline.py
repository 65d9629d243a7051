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
dbgs = [torch.zeros((148, 64), dtype=torch.int64, device=dev) for _ in range(iters)]
ext.lib.exl3b_debug_tc_timeline.argtypes = [ctypes.c_void_p]
path = int(os.environ.get("EXL3B_PATH", "0"))
ext.set_gemm_path(path)
# a CUDA graph of the back-to-back launches, like the benchmark
for i in range(2):
    ext.exl3_gemm(A, Bs[i], C, su, Ah, sv, -1, False, True, 0)
torch.cuda.synchronize()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for i in range(iters):
            ext.lib.exl3b_debug_tc_timeline(dbgs[i].data_ptr())
            ext.exl3_gemm(A, Bs[i], C, su, Ah, sv, -1, False, True, 0)
    ext.lib.exl3b_debug_tc_timeline(None)
    gr.replay(); torch.cuda.synchronize()
    gr.replay()
torch.cuda.synchronize()
names = {0: "entry", 1: "setup done", 2: "prologue done", 3: "decode: first W_FULL", 4: "decode: first unit done",
         5: "mma: first operands ready", 6: "mma: last commit", 10: "decode: last unit done", 7: "epi: first D_FULL",
         8: "epi: done", 11: "epi: partial written+fence", 12: "epi: counter known", 13: "epi(last): partials summed"}
D = [d.cpu().numpy().astype(np.int64) for d in dbgs]
t0 = min(d[d[:, 0] > 0][:, 0].min() for d in D)
print(f"knob={knob} path={path} shape k={k} n={n} K={K} m={m}: {iters} graph-replayed launches; ns since the first entry of launch 0 (min / median / max over CTAs)")
for i, d in enumerate(D):
    d = d[d[:, 0] > 0]
    print(f" launch {i}")
    for s in (0, 1, 2, 3, 4, 5, 10, 6, 7, 11, 12, 13, 8):
        v = d[:, s]; v = v[v > 0] - t0
        if len(v):
            print(f"  {names[s]:28s} {v.min():8d} {int(np.median(v)):8d} {v.max():8d}")

# int8 kernel: decode warp 4 at units 8 and 10, MMA warp at units 8..11
d = D[3]
if path == 210 and (d[:, 16] > 0).any():
    lab = ["loop top", "W_FULL", "LDS+SHFL", "A_EMPTY", "decode+STTM issued", "wait::st", "arrived"]
    base = d[:, 16]
    ok = base > 0
    for uu in range(2):
        row = [int(np.median((d[:, 16 + uu * 8 + i] - base)[ok & (d[:, 16 + uu * 8 + i] > 0)])) for i in range(7)]
        print("  i8 decode warp, unit", 8 + 2 * uu, dict(zip(lab, row)))
    units = (k // 128) * (n // 128) / 148
    wl = {56: "producer W_EMPTY", 57: "mma X_FULL", 58: "mma A_FULL", 59: "decode(g0) W_FULL", 60: "decode(g0) A_EMPTY", 61: "xf(w0) W_EMPTY"}
    print(f"  waits that found the barrier incomplete, median per CTA ({units:.1f} units/CTA; decode group / xf warp see half):",
          {v: int(np.median(d[:, s_][ok])) for s_, v in wl.items()})
    labm = ["loop top", "X_FULL", "A_FULL", "issued+committed"]
    for uu in range(4):
        row = [int(np.median((d[:, 32 + uu * 4 + i] - base)[ok & (d[:, 32 + uu * 4 + i] > 0)])) for i in range(4)]
        print("  i8 mma warp, unit", 8 + uu, dict(zip(labm, row)))
    sys.exit(0)

# fine-grained decode-warp stamps (exact kernel only): units 8..11 of each CTA, warp 4 lane 0
d = D[3]
if (d[:, 16] > 0).any():
    lab = ["loop top", "after W_FULL wait", "after LDS+SHFL", "after A_EMPTY wait", "after decode+STTM issue", "after wait::st", "after arrive"]
    print(" decode warp 4, launch 3: median ns since loop top of unit 8")
    base = d[:, 16]
    for uu in range(4):
        row = []
        for i in range(7):
            v = d[:, 16 + uu * 8 + i] - base
            v = v[d[:, 16 + uu * 8 + i] > 0]
            row.append(int(np.median(v)) if len(v) else -1)
        print("  unit", 8 + uu, dict(zip(lab, row)))
