"""
Host-side Hadamard helpers used by LinearEXL3.get_weight_tensor -- restated from the reference's
util/hadamard.py:34-42,107-131 (Sylvester construction) and modules/quant/exl3_lib/quantize.py:340-357
(preapply_had_l / preapply_had_r).  Runs on whatever device the tensor lives on (torch plumbing; not a hot path).
"""
from __future__ import annotations
import math
from functools import lru_cache
import torch


@lru_cache(maxsize=16)
def get_hadamard(n: int) -> torch.Tensor:
    assert n & (n - 1) == 0, "only power-of-two (Sylvester) sizes are needed by the EXL3 path"
    h = torch.ones((1, 1), dtype=torch.half)
    while h.shape[0] < n:
        h = torch.cat((torch.cat((h, h), 1), torch.cat((h, -h), 1)), 0)
    return h


@lru_cache(maxsize=16)
def get_hadamard_dt(n: int, device, dtype, scale=1.0) -> torch.Tensor:
    had = get_hadamard(n).to(device=device, dtype=dtype, copy=True)
    had *= scale
    return had


def preapply_had_l(x: torch.Tensor, had_dim: int) -> torch.Tensor:
    k, n = x.shape
    x_dtype = x.dtype
    x = x.to(torch.float)
    had = get_hadamard_dt(had_dim, x.device, x.dtype, 1 / math.sqrt(had_dim))
    x = (had @ x.view(-1, had_dim, n)).view(k, n)
    return x.to(x_dtype)


def preapply_had_r(x: torch.Tensor, had_dim: int) -> torch.Tensor:
    k, n = x.shape
    x_dtype = x.dtype
    x = x.to(torch.float)
    had = get_hadamard_dt(had_dim, x.device, x.dtype, 1 / math.sqrt(had_dim))
    x = (x.view(k, -1, had_dim) @ had).view(k, n)
    return x.to(x_dtype)
