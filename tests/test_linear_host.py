"""
Host logic of the LinearEXL3 mirror (exllamav3_b200/linear_exl3.py) against the reference's dispatch rules
(exllamav3/modules/quant/exl3.py:114-218), with a recording stand-in for the extension -- no kernel runs:
  * forward: rows <= 144 (or no_reconstruct) -> BC_LinearEXL3.run_alloc (the fused qgemm kernel); otherwise, or with
    params["reconstruct"], the reconstruct -> dense GEMM sibling; per-call override table; contiguity assertion; output dtype
  * reconstruct_hgemm: unfused sequence below 1024 rows (had_r_128 -> reconstruct -> hgemm -> had_r_128), fused above
    (reconstruct_had_slice -> hgemm), 32768-column slices for wide outputs with the svh slice and n_offset the kernels expect
"""
import numpy as np
import pytest
import torch
from exllamav3_b200 import ext, LinearEXL3
from exllamav3_b200 import linear_exl3 as le


class Rec:
    def __init__(self, mp):
        self.calls = []
        for name in ("had_r_128", "reconstruct", "reconstruct_slice", "reconstruct_had_slice", "hgemm", "exl3_gemm"):
            mp.setattr(ext, name, self._mk(name))

    def _mk(self, name):
        def f(*a):
            self.calls.append((name,) + tuple(tuple(t.shape) if isinstance(t, torch.Tensor) else t for t in a))
            return 210
        return f

    def names(self):
        return [c[0] for c in self.calls]


def _lin(k, n, K=4, **kw):
    g = torch.Generator().manual_seed(0)
    tr = torch.randint(0, 32767, (k // 16, n // 16, 16 * K), generator=g, dtype=torch.int32).to(torch.int16)
    return LinearEXL3(None, k, n, suh=torch.ones(k, dtype=torch.half), svh=torch.ones(n, dtype=torch.half), trellis=tr,
                      mul1=torch.zeros((), dtype=torch.int), key="blk.q_proj", **kw)


def test_forward_dispatch_rules(monkeypatch):
    rec = Rec(monkeypatch)
    lin = _lin(256, 384)
    y = lin.forward(torch.zeros((144, 256), dtype=torch.half), {})
    assert rec.names() == ["exl3_gemm"] and y.shape == (144, 384) and y.dtype == torch.half            # kernel path
    rec.calls.clear()
    y = lin.forward(torch.zeros((2, 72, 256), dtype=torch.half), {}, torch.float)                       # rows = product of leading dims
    assert rec.names() == ["exl3_gemm"] and y.shape == (2, 72, 384) and y.dtype == torch.float
    rec.calls.clear()
    lin.forward(torch.zeros((145, 256), dtype=torch.half), {})
    assert rec.names() == ["had_r_128", "reconstruct", "hgemm", "had_r_128"]                            # unfused sibling
    rec.calls.clear()
    lin.forward(torch.zeros((1, 256), dtype=torch.half), {"reconstruct": True})
    assert rec.names() == ["had_r_128", "reconstruct", "hgemm", "had_r_128"]
    rec.calls.clear()
    lin.config.infer_params.no_reconstruct = True
    lin.forward(torch.zeros((4096, 256), dtype=torch.half), {})
    assert rec.names() == ["exl3_gemm"]
    lin.config.infer_params.no_reconstruct = False
    with pytest.raises(AssertionError, match="non-contiguous"):
        lin.forward(torch.zeros((4, 512), dtype=torch.half)[:, ::2], {})
    # single rows go through the shared (1, k) scratch, larger inputs get a fresh one (linear.cpp:34-71)
    rec.calls.clear()
    lin.forward(torch.zeros((1, 256), dtype=torch.half), {}); lin.forward(torch.zeros((3, 256), dtype=torch.half), {})
    assert rec.calls[0][5] == (1, 256) and rec.calls[1][5] == (3, 256)                                  # A_had argument
    # per-call override table
    class Other:
        inner = None
        def forward(self, x, params, out_dtype=None): return "overridden"
    assert lin.forward(torch.zeros((1, 256), dtype=torch.half), {"ovr": {"blk.q_proj": Other()}}) == "overridden"


def test_reconstruct_hgemm_sequences(monkeypatch):
    rec = Rec(monkeypatch)
    lin = _lin(256, 384)
    lin.forward(torch.zeros((1024, 256), dtype=torch.half), {})
    assert rec.names() == ["reconstruct_had_slice", "hgemm"]                                            # fused above 1024 rows
    assert rec.calls[0][1] == (256, 384) and rec.calls[0][-1] == 0
    # wide outputs (lm_head): slices of at most 32768 columns, svh pre-offset, n_offset passed on
    monkeypatch.setattr(le, "MAX_RECONSTRUCT_SLICE_N", 256)
    wide = _lin(128, 640)
    rec.calls.clear()
    wide.forward(torch.zeros((2048, 128), dtype=torch.half), {})
    assert rec.names() == ["reconstruct_had_slice", "hgemm"] * 3
    offs = [c[-1] for c in rec.calls if c[0] == "reconstruct_had_slice"]
    widths = [c[1][1] for c in rec.calls if c[0] == "reconstruct_had_slice"]
    svh_len = [c[4][0] for c in rec.calls if c[0] == "reconstruct_had_slice"]
    assert offs == [0, 256, 512] and widths == [256, 256, 128] and svh_len == [640, 384, 128]
    outs = [c[3] for c in rec.calls if c[0] == "hgemm"]
    assert outs == [(2048, 256), (2048, 256), (2048, 128)]
    rec.calls.clear()
    wide.forward(torch.zeros((200, 128), dtype=torch.half), {})                                         # unfused + sliced
    assert rec.names() == ["had_r_128"] + ["reconstruct_slice", "hgemm"] * 3 + ["had_r_128"]


def test_weight_tensor_composition(monkeypatch):
    """get_weight_tensor = diag(suh) H128 W_hat H128 diag(svh) (exl3.py:227-237) from the inner weights."""
    k, n = 128, 256
    lin = _lin(k, n)
    rng = np.random.default_rng(0)
    W = torch.from_numpy(rng.standard_normal((k, n)).astype(np.float16))
    lin.suh = torch.from_numpy(rng.standard_normal(k).astype(np.float16)); lin.svh = torch.from_numpy(rng.standard_normal(n).astype(np.float16))
    monkeypatch.setattr(ext, "reconstruct", lambda w, tr, K, mcg, mul1: w.copy_(W))
    got = lin.get_weight_tensor().double()
    from oracle import exl3_oracle as orc
    H = torch.from_numpy(orc.hadamard_matrix_128() / np.sqrt(128.0))
    want = (H @ W.double()) * lin.suh.double().unsqueeze(1)
    want = torch.cat([want[:, i:i + 128] @ H for i in range(0, n, 128)], dim=1) * lin.svh.double().unsqueeze(0)
    assert float((got - want).abs().max() / want.abs().max()) < 5e-3
