"""reconstruct_had: CUDA-core kernel (mode 1) against the tcgen05 kernel (mode 2), us and GB/s per shape (graph replay, L2 flushed by size)."""
import json, sys, torch
sys.path.insert(0, ".")
from exllamav3_b200 import ext

dev = torch.device("cuda:0")
shapes = [(4096, 4096, 4), (4096, 14336, 4), (14336, 4096, 4), (8192, 8192, 4), (4096, 14336, 2), (4096, 14336, 6), (4096, 4096, 8)]
for (k, n, K) in shapes:
    g = torch.Generator(device=dev); g.manual_seed(1)
    nrep = max(2, int(300e6 // (k * n * (2 + K / 8))) + 1)      # distinct tensors: working set beyond L2
    trs = [torch.randint(-32768, 32767, (k // 16, n // 16, 16 * K), dtype=torch.int16, device=dev, generator=g) for _ in range(nrep)]
    suh = (torch.randn(k, device=dev, generator=g) / k ** 0.5).half(); svh = torch.randn(n, device=dev, generator=g).half()
    ws = [torch.empty((k, n), dtype=torch.half, device=dev) for _ in range(nrep)]
    row = {"k": k, "n": n, "K": K}
    for mode, nm in ((1, "cuda_core"), (21, "tensor_core_x1"), (22, "tensor_core_x2"), (24, "tensor_core_x4")):
        ext.lib.exl3b_debug_reconstruct_had(mode)
        for i in range(nrep): ext.reconstruct_had_slice(ws[i], trs[i], suh, svh, K, False, True, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            for i in range(nrep): ext.reconstruct_had_slice(ws[i], trs[i], suh, svh, K, False, True, 0)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (iters * nrep)
        b = k * n * (2 + K / 8)
        row[nm] = {"us": round(us, 2), "GBps": round(b / us / 1e3, 1)}
    ext.lib.exl3b_debug_reconstruct_had(0)
    print(json.dumps(row), flush=True)
