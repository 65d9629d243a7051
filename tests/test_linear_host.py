"""
Host logic of QLinear (exllamav3_b200/qlinear.py) against the dispatch rules of the path (the reference's
exllamav3/modules/quant/exl3.py:114-218 defines them), with a recording stand-in for the extension -- no kernel runs:
  * forward: rows <= 144 -> BC_LinearEXL3.run_alloc (the fused qgemm kernel); otherwise, or with params["reconstruct"], the
    materialise -> dense GEMM sibling; contiguity check; output dtype
  * forward_dense: unfolded sequence below 1024 rows (had_r_128 -> reconstruct -> hgemm -> had_r_128), folded above
    (reconstruct_had_slice -> hgemm), 32768-column windows for wide outputs with the svh slice and n_offset the kernels expect
"""
import numpy as np
import pytest
import torch
from exllamav3_b200 import ext, QLinear
from exllamav3_b200 import qlinear as le


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
    return QLinear(tr, torch.ones(k, dtype=torch.half), torch.ones(n, dtype=torch.half), mul1=True, **kw)


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
    lin.forward(torch.zeros((4096, 256), dtype=torch.half), {"no_reconstruct": True})
    assert rec.names() == ["exl3_gemm"]
    with pytest.raises(ValueError, match="non-contiguous"):
        lin.forward(torch.zeros((4, 512), dtype=torch.half)[:, ::2], {})
    # single rows go through the shared (1, k) scratch, larger inputs get a fresh one (linear.cpp:34-71)
    rec.calls.clear()
    lin.forward(torch.zeros((1, 256), dtype=torch.half), {}); lin.forward(torch.zeros((3, 256), dtype=torch.half), {})
    assert rec.calls[0][5] == (1, 256) and rec.calls[1][5] == (3, 256)                                  # A_had argument
    with pytest.raises(ValueError, match="int16"):
        QLinear(torch.zeros((16, 16, 64), dtype=torch.int32), lin.suh, lin.svh)
    with pytest.raises(ValueError, match="do not match"):
        QLinear(lin.trellis, lin.suh[:128], lin.svh)


def test_reconstruct_hgemm_sequences(monkeypatch):
    rec = Rec(monkeypatch)
    lin = _lin(256, 384)
    lin.forward(torch.zeros((1024, 256), dtype=torch.half), {})
    assert rec.names() == ["reconstruct_had_slice", "hgemm"]                                            # fused above 1024 rows
    assert rec.calls[0][1] == (256, 384) and rec.calls[0][-1] == 0
    # wide outputs (lm_head): slices of at most 32768 columns, svh pre-offset, n_offset passed on
    monkeypatch.setattr(le, "DENSE_WINDOW_COLS", 256)
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


def test_weight_accessors_use_the_reconstruct_kernels(monkeypatch):
    """weight_inner = decoded trellis values (reconstruct); weight = original-basis weights through the fused kernel
    (reconstruct_had_slice with the full svh and offset 0) -- no torch-side Hadamard anywhere in the product."""
    rec = Rec(monkeypatch)
    lin = _lin(128, 256)
    w = lin.weight_inner()
    assert rec.names() == ["reconstruct"] and w.shape == (128, 256) and w.dtype == torch.half
    rec.calls.clear()
    lin.weight()
    assert rec.names() == ["reconstruct_had_slice"] and rec.calls[0][1] == (128, 256) and rec.calls[0][4] == (256,) and rec.calls[0][-1] == 0


def test_pointer_tables_validate_shapes():
    from exllamav3_b200 import pointer_tables
    a, b, c = _lin(128, 256), _lin(128, 256), _lin(256, 256)
    B, su, sv = pointer_tables("cpu", [a, b])
    assert B.tolist() == [a.trellis.data_ptr(), b.trellis.data_ptr()] and su.dtype == sv.dtype == torch.long
    with pytest.raises(ValueError, match="identical shapes"):
        pointer_tables("cpu", [a, c])


def test_use_mgemm_answer_of_the_shim():
    import inspect
    src = inspect.getsource(ext.exl3_gemv_int8_max_k)
    assert "return 0" in src
