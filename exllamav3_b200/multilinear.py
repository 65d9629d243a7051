"""
MultiLinear -- device pointer tables (trellis / suh / svh addresses) for exl3_mgemm, mirror of the reference's
modules/multilinear.py:5-43.  Accepts LinearEXL3 objects directly (the reference wraps them in `Linear`, whose
`.inner` is the LinearEXL3).
"""
from __future__ import annotations
import torch


def _inner(l):
    return getattr(l, "inner", l)


class MultiLinear:
    def __init__(self, device, linears: list, allow_bias: bool = False):
        self.device = device
        self.linears = linears
        self.num_linears = len(linears)
        inner = [_inner(l) for l in linears]
        assert all(i.quant_type == "exl3" for i in inner)
        assert allow_bias or all(i.bias is None for i in inner)
        self.in_features = inner[0].in_features
        self.out_features = inner[0].out_features
        self.K = inner[0].K
        assert all(i.K == self.K for i in inner)
        assert all(i.in_features == self.in_features for i in inner)
        assert all(i.out_features == self.out_features for i in inner)
        self.ptrs_suh = torch.tensor([i.suh.data_ptr() for i in inner], dtype=torch.long, device=device)
        self.ptrs_svh = torch.tensor([i.svh.data_ptr() for i in inner], dtype=torch.long, device=device)
        self.ptrs_trellis = torch.tensor([i.trellis.data_ptr() for i in inner], dtype=torch.long, device=device)
        self.mcg = inner[0].mcg
        assert all(i.mcg == self.mcg for i in inner[1:])
        self.mul1 = inner[0].mul1
        assert all(i.mul1 == self.mul1 for i in inner[1:])

    def q_cb(self):
        return self.mcg, self.mul1

    def unload(self):
        pass
