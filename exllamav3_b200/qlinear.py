"""
QLinear -- the smallest useful caller of the operator surface in `ext`: one EXL3 tensor set plus the two ways of applying it.

This is NOT the reference's `LinearEXL3` (modules/quant/exl3.py): that class is meant to run on the shim unchanged
(tests/test_reference_callsites.py drives the reference's own file over `ext` where the checkout exists).
QLinear exists so that bench.py, smoke() and the GPU tests have a caller on boxes without the reference checkout.  What it
shares with the reference is the dispatch RULE, which is part of the path's behaviour:

    rows <= 144            the fused decode-GEMM            ext.BC_LinearEXL3.run_alloc -> exl3_gemm
    rows  > 144 / forced   weights materialised, dense GEMM  reconstruct* -> hgemm, in column windows of <= 32768;
                           from 1024 rows on the Hadamards and scales are folded into the materialised weights
                           (reconstruct_had_slice), below that they are applied to the activations (had_r_128)
    (thresholds: modules/quant/exl3.py:10-12,176)
"""
from __future__ import annotations
import torch
from . import ext

KERNEL_MAX_ROWS = 144            # above this the dense sibling is faster (reference: AUTO_RECONSTRUCT_THRESHOLD)
DENSE_WINDOW_COLS = 32768        # widest weight window materialised at once (reference: MAX_RECONSTRUCT_SLICE_N)
FOLDED_MIN_ROWS = 1024           # fold Hadamards + scales into the window from here on

_row_scratch: dict = {}


def _scratch_row(device, k: int) -> torch.Tensor:
    """One (1, k) fp16 row per device and width, shared by every QLinear of that width (the reference shares the same
    scratch through its process-global tensor cache; single-stream use only)."""
    key = (str(device), k)
    if key not in _row_scratch:
        _row_scratch[key] = torch.empty((1, k), dtype=torch.half, device=device)
    return _row_scratch[key]


class QLinear:
    def __init__(self, trellis: torch.Tensor, suh: torch.Tensor, svh: torch.Tensor, *, mcg: bool = False, mul1: bool = False,
                 bias: torch.Tensor | None = None, out_dtype: torch.dtype = torch.half, bias_in_group: bool | None = None):
        if trellis.dtype != torch.int16 or trellis.dim() != 3 or trellis.shape[2] % 16:
            raise ValueError("trellis must be int16 of shape (k/16, n/16, 16*K)")
        if suh.dtype != torch.half or svh.dtype != torch.half:
            raise ValueError("suh / svh must be fp16 (unpack legacy sign bitfields before constructing a QLinear)")
        self.trellis, self.suh, self.svh = trellis, suh, svh
        self.in_features, self.out_features, self.K = trellis.shape[0] * 16, trellis.shape[1] * 16, trellis.shape[2] // 16
        if suh.numel() != self.in_features or svh.numel() != self.out_features:
            raise ValueError("suh / svh do not match the trellis shape")
        self.mcg, self.mul1 = bool(mcg), bool(mul1)
        self.bias = bias.to(torch.half) if bias is not None and bias.dtype == torch.float else bias
        self.out_dtype = out_dtype
        # tensor parallel: does ANY shard of this linear carry a bias?  Every rank must agree on the answer, because it
        # selects the collective (tp.row_parallel_forward): recorded on all shards by tp.tp_slice.
        self.bias_in_group = (bias is not None) if bias_in_group is None else bool(bias_in_group)
        self.op = ext.BC_LinearEXL3(trellis, suh, svh, self.K, self.bias, self.mcg, self.mul1,
                                    _scratch_row(trellis.device, self.in_features))

    # ---- apply ---------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, params: dict | None = None, out_dtype: torch.dtype | None = None) -> torch.Tensor:
        if not x.is_contiguous():
            raise ValueError(f"QLinear: non-contiguous input {tuple(x.shape)}")
        rows = x.numel() // x.shape[-1] if x.shape[-1] else 0
        dense = bool(params and params.get("reconstruct")) or rows > KERNEL_MAX_ROWS
        if params and params.get("no_reconstruct"):
            dense = False
        dt = out_dtype or self.out_dtype
        if not dense:
            return self.op.run_alloc(x, self.out_features, dt == torch.float)
        return self.forward_dense(x, dt)

    __call__ = forward

    def _window(self, c0: int, c1: int, folded: bool, buf: torch.Tensor | None = None) -> torch.Tensor:
        """fp16 weights of output columns [c0, c1): trellis values (folded = False) or original-basis weights."""
        n = c1 - c0
        w = (buf[: self.in_features * n].view(self.in_features, n) if buf is not None
             else torch.empty((self.in_features, n), dtype=torch.half, device=self.trellis.device))
        if folded:
            ext.reconstruct_had_slice(w, self.trellis, self.suh, self.svh[c0:], self.K, self.mcg, self.mul1, c0)
        elif c0 == 0 and c1 == self.out_features:
            ext.reconstruct(w, self.trellis, self.K, self.mcg, self.mul1)
        else:
            ext.reconstruct_slice(w, self.trellis, self.K, self.mcg, self.mul1, c0)
        return w

    def forward_dense(self, x: torch.Tensor, out_dtype: torch.dtype | None = None) -> torch.Tensor:
        k, n = self.in_features, self.out_features
        x2 = x.view(-1, k)
        rows = x2.shape[0]
        y = torch.empty(tuple(x.shape[:-1]) + (n,), dtype=out_dtype or self.out_dtype, device=x.device)
        y2 = y.view(rows, n)
        folded = rows >= FOLDED_MIN_ROWS and k % 128 == 0 and n % 128 == 0
        a = x2
        if not folded:
            a = torch.empty_like(x2)
            ext.had_r_128(x2, a, self.suh, None, 1.0)
        windows = [(c, min(c + DENSE_WINDOW_COLS, n)) for c in range(0, n, DENSE_WINDOW_COLS)]
        buf = (torch.empty((k * DENSE_WINDOW_COLS,), dtype=torch.half, device=self.trellis.device) if len(windows) > 1 else None)
        for c0, c1 in windows:
            ext.hgemm(a, self._window(c0, c1, folded, buf), y2[:, c0:c1])
        if not folded:
            ext.had_r_128(y2, y2, None, self.svh, 1.0)
        if self.bias is not None:
            y += self.bias
        return y

    # ---- weights ---------------------------------------------------------------------------------------------------
    def weight_inner(self) -> torch.Tensor:
        """Decoded trellis values (k, n) fp16, before any Hadamard or scale."""
        return self._window(0, self.out_features, False)

    def weight(self) -> torch.Tensor:
        """Original-basis weights  diag(suh) H W_hat H diag(svh)  (k, n) fp16, through the fused reconstruct kernel."""
        return self._window(0, self.out_features, True)


def pointer_tables(device, linears: list[QLinear]):
    """(B_ptrs, suh_ptrs, svh_ptrs) int64 device tensors for exl3_mgemm over same-shape, same-codebook linears."""
    a = linears[0]
    for l in linears:
        if (l.in_features, l.out_features, l.K, l.mcg, l.mul1) != (a.in_features, a.out_features, a.K, a.mcg, a.mul1):
            raise ValueError("exl3_mgemm needs identical shapes, bitrates and codebooks")
    mk = lambda f: torch.tensor([f(l).data_ptr() for l in linears], dtype=torch.long, device=device)
    return mk(lambda l: l.trellis), mk(lambda l: l.suh), mk(lambda l: l.svh)
