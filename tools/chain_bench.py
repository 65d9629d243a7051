"""Timing of the chain kernel against the single-GEMM int8 kernel (not part of the product).
   python tools/chain_bench.py            per-shape us (graph replay over distinct weights), block / layer / token chains of Llama-3.1-8B"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from exllamav3_b200 import ext

dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)


def mk(k, n, K, c_fp32=True):
    tr = torch.randint(0, 65536, (k // 16, n // 16, 16 * K), generator=g, device=dev, dtype=torch.int32).to(torch.int16)
    sgn = lambda sz: (torch.randint(0, 2, (sz,), generator=g, device=dev) * 2 - 1).float()
    suh = (sgn(k) * (0.5 + 1.5 * torch.rand(k, generator=g, device=dev)) / k ** 0.5).half()
    svh = (sgn(n) * (0.5 + 1.5 * torch.rand(n, generator=g, device=dev))).half()
    x = torch.randn((1, k), generator=g, device=dev).half()
    y = torch.empty((1, n), dtype=torch.float if c_fp32 else torch.half, device=dev)
    return dict(tr=tr, suh=suh, svh=svh, x=x, y=y, xh=torch.empty_like(x), k=k, n=n, K=K)


def timed(fn, reps=5, inner=1):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); fn()
    st.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        fn()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(reps):
                gr.replay()
            e1.record(st)
        e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps / inner
        best = us if best is None else min(best, us)
    return best


def alg_bytes(k, n, K, c_fp32=True):
    return k * n * K // 8 + 2 * k + n * (4 if c_fp32 else 2) + 2 * (k + n)


out = {}
copies = 16
for name, (k, n, K) in dict(q=(4096, 4096, 4), kv=(4096, 1024, 4), gate=(4096, 14336, 4), down=(14336, 4096, 4), head=(4096, 128256, 6),
                            q_K2=(4096, 4096, 2), q_K3=(4096, 4096, 3), q_K6=(4096, 4096, 6)).items():
    mats = [mk(k, n, K) for _ in range(copies if n < 100000 else 2)]
    res = {}
    for path in (ext.EXL3B_TAG_TC_I8, ext.EXL3B_TAG_TC_I8_CHAIN):
        ext.set_gemm_path(path)
        def run():
            for t in mats:
                ext.exl3_gemm(t["x"], t["tr"], t["y"], t["suh"], t["xh"], t["svh"], -1, False, True, 0)
        res[path] = timed(run, inner=len(mats))
    ext.set_gemm_path(0)
    # the same GEMMs as ONE chain with a dependency between consecutive ops (what a dependent sequence costs inside one launch)
    ch = ext.GemmChain([dict(x=t["x"], trellis=t["tr"], suh=t["suh"], svh=t["svh"], y=t["y"], mul1=True, new_stage=i > 0) for i, t in enumerate(mats)])
    res["chain_dep"] = timed(ch.run, inner=len(mats))
    b = alg_bytes(k, n, K)
    out[name] = {"k": k, "n": n, "K": K, "us_v1": round(res[ext.EXL3B_TAG_TC_I8], 2), "us_chain_single": round(res[ext.EXL3B_TAG_TC_I8_CHAIN], 2),
                 "us_chain_dependent_ops": round(res["chain_dep"], 2), "GBps_chain_dep": round(b / res["chain_dep"] / 1e3)}
    print(name, out[name], flush=True)
    del mats, ch

# Llama-3.1-8B token three ways
layers = 32
L = [dict(q=mk(4096, 4096, 4, False), k=mk(4096, 1024, 4, False), v=mk(4096, 1024, 4, False), o=mk(4096, 4096, 4), gate=mk(4096, 14336, 4),
          up=mk(4096, 14336, 4), down=mk(14336, 4096, 4)) for _ in range(layers)]
head = mk(4096, 128256, 6)
op = lambda t, **kw: dict(trellis=t["tr"], suh=t["suh"], svh=t["svh"], y=t["y"], mul1=True, **kw)


def layer_ops(l, first_new):
    return [op(l["q"], x=l["q"]["x"], new_stage=first_new), op(l["k"], x=l["q"]["x"]), op(l["v"], x=l["q"]["x"]),
            op(l["o"], x=l["q"]["y"], new_stage=True),
            op(l["gate"], x=l["gate"]["x"], new_stage=True), op(l["up"], x=l["gate"]["x"]),
            op(l["down"], gate=l["gate"]["y"], up=l["up"]["y"], new_stage=True)]


tok = {}
blocks = []
for l in L:
    ops = layer_ops(l, False)
    blocks += [ext.GemmChain(ops[0:3]), ext.GemmChain([dict(ops[3], new_stage=False)]), ext.GemmChain([dict(ops[4], new_stage=False), ops[5], ops[6]])]
blocks.append(ext.GemmChain([op(head, x=head["x"])]))
tok["block_chains_97_launches"] = timed(lambda: [c.run() for c in blocks])
per_layer = [ext.GemmChain(layer_ops(l, False)) for l in L] + [blocks[-1]]
tok["layer_chains_33_launches"] = timed(lambda: [c.run() for c in per_layer])
allops = []
for i, l in enumerate(L):
    allops += layer_ops(l, i > 0)
allops.append(op(head, x=head["x"], new_stage=True))
whole = ext.GemmChain(allops)
tok["token_chain_1_launch"] = timed(whole.run)
total_b = sum(alg_bytes(t["k"], t["n"], t["K"], t["y"].dtype == torch.float) for l in L for t in l.values()) + alg_bytes(4096, 128256, 6)
for kname, us in tok.items():
    print(kname, round(us / 1e3, 3), "ms/token", round(1e6 / us, 1), "tok/s", round(total_b / us / 1e3), "GB/s", flush=True)
out["token"] = {kname: {"ms": round(us / 1e3, 4), "tok_s": round(1e6 / us, 1), "GBps": round(total_b / us / 1e3)} for kname, us in tok.items()}
print(json.dumps(out))
