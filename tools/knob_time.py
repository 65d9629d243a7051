"""Where does the time per 128x128 unit go?  Times the int8 tensor-core decode-GEMM with parts of the pipeline switched
off (debug build only: EXL3B_TC_DEBUG=1 python exllamav3_b200/build.py -f).  Results are wrong by construction.
   knob bits: 1 skip decode math, 2 skip tcgen05.st, 4 skip MMA, 8 skip activation digits."""
import os, sys, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from exllamav3_b200 import ext

lib = ext._lib
dev = torch.device("cuda:0")
path = int(os.environ.get("EXL3B_PATH", "210"))
ext.set_gemm_path(path)
shapes = [(4096, 128256, 6, 2), (4096, 14336, 4, 8), (4096, 4096, 4, 24), (14336, 4096, 4, 8)]
knobs = [int(v) for v in os.environ.get('KNOBS', '0,1,3,4,7,8,15').split(',')]
for (k, n, K, copies) in shapes:
    g = torch.Generator(device=dev); g.manual_seed(1)
    trs = [torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(copies)]
    suh = torch.ones(k, dtype=torch.half, device=dev); svh = torch.ones(n, dtype=torch.half, device=dev)
    M = int(os.environ.get("M", "1"))
    x = torch.randn((M, k), generator=g, device=dev).half(); xh = torch.empty_like(x)
    y = torch.empty((M, n), dtype=torch.float, device=dev)
    row = {}
    for kn in knobs:
        lib.exl3b_debug_tc_knob(kn)
        def run():
            for tr in trs:
                ext.exl3_gemm(x, tr, y, suh, xh, svh, -1, False, True, 0)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(); run()
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            with torch.cuda.stream(s):
                e0.record(s)
                for _ in range(4): gr.replay()
                e1.record(s)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / (4 * copies) * 1e3)
        row[kn] = round(best, 2)
    lib.exl3b_debug_tc_knob(0)
    units = (k // 128) * (n // 128) / 148
    print(json.dumps({"k": k, "n": n, "K": K, "units_per_cta": round(units, 1), "us_by_knob": row,
                      "GBps_knob0": round(k * n * K / 8 / row[0] / 1e3)}), flush=True)
