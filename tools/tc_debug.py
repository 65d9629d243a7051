"""Bring-up diagnostics for the tcgen05 path (not a test): each case runs in its own process so a trap/hang in one
does not poison the others.   python tools/tc_debug.py [case...]"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (m, k, n, K, cb, transforms, fp32, onehot_k)
    "unit_onehot0":  (1, 128, 128, 4, 2, False, True, 0),
    "unit_onehot37": (1, 128, 128, 4, 2, False, True, 37),
    "unit_raw":      (1, 128, 128, 4, 2, False, True, None),
    "unit_full":     (1, 128, 128, 4, 2, True, True, None),
    "k512":          (1, 512, 128, 4, 2, True, True, None),
    "n512":          (1, 128, 512, 4, 2, True, True, None),
    "sq1024":        (1, 1024, 1024, 4, 2, True, False, None),
    "m5":            (5, 512, 384, 4, 2, True, True, None),
    "m16":           (16, 512, 384, 4, 2, True, False, None),
    "m17":           (17, 512, 384, 4, 2, True, True, None),
    "m100":          (100, 256, 256, 4, 2, True, True, None),
    "K3cb0":         (3, 512, 384, 3, 0, True, True, None),
    "K6cb1":         (2, 512, 384, 6, 1, True, True, None),
    "K1":            (1, 256, 256, 1, 2, True, True, None),
    "K8":            (1, 256, 256, 8, 2, True, True, None),
    "big":           (1, 4096, 4096, 4, 2, True, False, None),
    "big_gate":      (1, 4096, 14336, 4, 2, True, True, None),
}


def run_case(name):
    import numpy as np, torch
    from exllamav3_b200 import ext
    from oracle import exl3_oracle as orc
    m, k, n, K, cb, tr_on, fp32, onehot = CASES[name]
    dev = torch.device("cuda:0")
    tr, suh, svh, x = orc.make_synthetic(k, n, K, m=m)
    if onehot is not None:
        x = np.zeros((m, k), np.float16); x[0, onehot] = 1.0
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ext.set_gemm_path(int(os.environ.get("EXL3B_PATH", "200")))
    C = torch.full((m, n), float("nan"), dtype=torch.float if fp32 else torch.half, device=dev)
    A = T(x)
    tag = ext.exl3_gemm(A, T(tr), C, T(suh) if tr_on else None, torch.empty_like(A) if tr_on else None,
                        T(svh) if tr_on else None, -1, cb == 1, cb == 2, 0)
    torch.cuda.synchronize()
    y = C.cpu().numpy().astype(np.float64)
    if tr_on:
        ref = orc.exl3_gemm_f64(x, tr, suh, svh, K, cb)
    else:
        ref = x.astype(np.float64) @ orc.reconstruct(tr, K, cb).astype(np.float64)
    err = np.abs(y - ref)
    out = dict(case=name, tag=tag, nan=int(np.isnan(y).sum()), max_rel=float(np.nanmax(err) / np.abs(ref).max()),
               rms_rel=float(np.sqrt(np.nanmean(err ** 2)) / np.sqrt((ref ** 2).mean())))
    if out["max_rel"] > 3e-3 or out["nan"]:
        bad = np.argwhere(~(err <= 3e-3 * np.abs(ref).max()))
        out["n_bad"] = int(len(bad)); out["first_bad"] = bad[:8].tolist()
        out["y0"] = [float(v) for v in y[0, :8]]; out["ref0"] = [float(v) for v in ref[0, :8]]
        if onehot is not None:
            W = orc.reconstruct(tr, K, cb).astype(np.float64)
            # which k-row does the output resemble, and under which column permutation?
            best = min(range(k), key=lambda kk: np.abs(np.sort(W[kk]) - np.sort(np.nan_to_num(y[0]))).sum())
            out["resembles_row"] = int(best)
            cols = [int(np.argmin(np.abs(W[best] - y[0, j]))) for j in range(16)]
            out["col_map_first16"] = cols
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
    else:
        names = sys.argv[1:] or list(CASES)
        for nme in names:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", nme], capture_output=True, text=True, timeout=180)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            print(nme, "rc", r.returncode, tail if tail else r.stderr[-600:], flush=True)
