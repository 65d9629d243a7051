"""Round-2 experiment: does a smaller persistent grid help the short launches?  exl3_gemm's force_num_sms caps the grid
(plan_gemm_tc_i8); with KB = k / 128 units per column strip, a grid that divides the unit count evenly puts every CTA in ONE
strip (4096 x 4096: 1024 units, 128 CTAs x 8 units, 4 CTAs per strip) instead of the ragged 148-way split -- fewer split-K
partners and a shorter tail per launch, against 13 % fewer SMs streaming.  Times CUDA-graph replays of the Llama-3.1-8B shapes
(m = 1, mul1) per grid size and prints one JSON line per shape.  Not part of the product.
   python tools/sweep_grid.py            [GRIDS=148,144,128,... M=1]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from exllamav3_b200 import ext

dev = torch.device("cuda:0")
grids = [int(v) for v in os.environ.get("GRIDS", "148,144,136,128,120,112,96,80,74,64").split(",")]
M = int(os.environ.get("M", "1"))
shapes = [(4096, 4096, 4, 24), (4096, 1024, 4, 48), (4096, 14336, 4, 8), (14336, 4096, 4, 8), (4096, 128256, 6, 2)]
for (k, n, K, copies) in shapes:
    g = torch.Generator(device=dev); g.manual_seed(1)
    trs = [torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(copies)]
    suh = (torch.randn(k, generator=g, device=dev) / k ** 0.5).half(); svh = torch.randn(n, generator=g, device=dev).half()
    x = torch.randn((M, k), generator=g, device=dev).half(); xh = torch.empty_like(x)
    y = torch.empty((M, n), dtype=torch.float, device=dev)
    row = {}
    for G in grids:
        def run():
            for tr in trs:
                ext.exl3_gemm(x, tr, y, suh, xh, svh, -1, False, True, G)
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
        row[G] = round(best, 2)
    units = (k // 128) * (n // 128)
    print(json.dumps({"k": k, "n": n, "K": K, "m": M, "units": units, "KB": k // 128, "us_by_grid": row,
                      "best_grid": min(row, key=row.get)}), flush=True)
