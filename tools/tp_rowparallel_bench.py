"""Round-2 measurement: one row-parallel output, two ways, per shape (run under torchrun on N GPUs of one box):
     exl3_gemm + NCCL all_reduce      (what the reference's TP mode issues, modules/mlp.py:769-770)
     exl3_gemm_allreduce              (ONE kernel: partials exchanged over NVLink peer memory in the epilogue)
   CUDA-graph replays of `copies` distinct weight shards back to back, CUDA events, max over ranks.  Not part of the product.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/tp_rowparallel_bench.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from exllamav3_b200 import ext, tp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tp.enable_fused_allreduce(max_elems=4 * 8192)
shapes = [("8b o", 4096, 4096), ("8b down", 14336, 4096), ("70b o", 8192, 8192), ("70b down", 28672, 8192)]
K, copies = 4, 8
for (name, k_full, n) in shapes:
    k = max(128, (k_full // world) // 128 * 128)
    g = torch.Generator(device=dev); g.manual_seed(1 + rank)
    trs = [torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16) for _ in range(copies)]
    suh = (torch.randn(k, generator=g, device=dev) / k_full ** 0.5).half(); svh = torch.randn(n, generator=g, device=dev).half()
    x = torch.randn((1, k), generator=g, device=dev).half(); xh = torch.empty_like(x)
    y = torch.empty((1, n), dtype=torch.float, device=dev)
    res = {}
    for mode in ("nccl", "fused"):
        def run():
            for tr in trs:
                if mode == "fused":
                    ext.exl3_gemm_allreduce(x, tr, y, suh, None, svh, False, True)
                else:
                    ext.exl3_gemm(x, tr, y, suh, xh, svh, -1, False, True, 0)
                    dist.all_reduce(y)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(); run()
        s.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            run()
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s):
                e0.record(s)
                for _ in range(4): gr.replay()
                e1.record(s)
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / (4 * copies) * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        res[mode] = round(best, 2)
        del gr
    if rank == 0:
        print(json.dumps({"shape": name, "k_shard": k, "n": n, "world": world, "us_per_output": res}), flush=True)
torch.cuda.synchronize(); dist.barrier()
tp.disable_fused_allreduce()
os._exit(0)
