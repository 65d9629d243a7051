"""Bring-up / timing of the tcgen05 dense GEMM behind ext.hgemm.  python tools/hgemm_debug.py [--time]"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [(128, 64, 256), (128, 128, 256), (1, 128, 128), (37, 256, 384), (300, 512, 256), (256, 4096, 512), (2048, 4096, 4096),
         (129, 72, 264), (4096, 14336, 4096)]


def one(m, k, n, fp32):
    import torch
    from exllamav3_b200 import ext
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(m + k + n)
    a = torch.randn((m, k), generator=g, device=dev).half(); b = torch.randn((k, n), generator=g, device=dev).half()
    c = torch.full((m, n), float("nan"), dtype=torch.float if fp32 else torch.half, device=dev)
    ext.hgemm(a, b, c)
    torch.cuda.synchronize()
    ref = a.float() @ b.float()
    err = (c.float() - ref).abs()
    bad = ~(err <= 2e-3 * ref.abs().max() + 1e-3)
    out = dict(m=m, k=k, n=n, fp32=fp32, nan=int(torch.isnan(c.float()).sum()), max_rel=float(err.nan_to_num(1e9).max() / ref.abs().max()),
               n_bad=int(bad.sum()))
    if out["n_bad"]:
        idx = bad.nonzero()[:6].tolist()
        out["first_bad"] = idx
        out["got"] = [float(c[i, j]) for i, j in idx[:4]]; out["want"] = [float(ref[i, j]) for i, j in idx[:4]]
        rows = sorted(set(i for i, _ in bad.nonzero().tolist()))[:10]; cols = sorted(set(j for _, j in bad.nonzero().tolist()))[:16]
        out["bad_rows"] = rows; out["bad_cols"] = cols
    print(json.dumps(out), flush=True)


def timeit():
    import torch
    from exllamav3_b200 import ext
    dev = torch.device("cuda:0")
    for (m, k, n) in [(2048, 4096, 4096), (8192, 4096, 4096), (65536, 4096, 4096), (8192, 4096, 14336), (8192, 14336, 4096)]:
        a = torch.randn((m, k), device=dev).half(); b = torch.randn((k, n), device=dev).half()
        c = torch.empty((m, n), dtype=torch.half, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 10
        res = {}
        for mode, name in ((1, "single_cta"), (2, "cta_pair")):
            ext.lib.exl3b_debug_hgemm_pair(mode)
            for _ in range(3): ext.hgemm(a, b, c)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(it): ext.hgemm(a, b, c)
            e1.record(); e1.synchronize()
            res[name] = e0.elapsed_time(e1) / it
        ext.lib.exl3b_debug_hgemm_pair(0)
        ms = min(res.values())
        tf = 2.0 * m * k * n / ms / 1e9
        for _ in range(3): torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(it): torch.matmul(a, b, out=c)
        e1.record(); e1.synchronize()
        ms2 = e0.elapsed_time(e1) / it
        print(json.dumps(dict(m=m, k=k, n=n, ms=ms, tflops=tf, ms_single_cta=res["single_cta"], ms_cta_pair=res["cta_pair"],
                              tflops_cta_pair=2.0 * m * k * n / res["cta_pair"] / 1e9, cublas_ms=ms2, cublas_tflops=2.0 * m * k * n / ms2 / 1e9,
                              pair_vs_cublas=ms2 / res["cta_pair"])), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] == "1")
    elif len(sys.argv) > 1 and sys.argv[1] == "--time":
        timeit()
    else:
        for (m, k, n) in CASES:
            for fp32 in (1, 0):
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(m), str(k), str(n), str(fp32)],
                                   capture_output=True, text=True, timeout=180)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                print("rc", r.returncode, tail if tail else r.stderr[-500:], flush=True)
