"""
Tensor-parallel shard + collective plumbing around the EXL3 linear -- the part of the reference's TP machinery that
touches the qgemm path:

  * shard construction: LinearEXL3.tp_import_split slicing rules (modules/quant/exl3.py:284-330) via LinearEXL3.tp_slice
  * partition: q/k/v/gate/up column-parallel, o/down row-parallel, 128-channel granularity (modules/linear.py:645-656)
  * one all-reduce (sum) per row-parallel output, called by the CALLER of the linear (modules/mlp.py:769-770,
    modules/attn.py:546-547) -- here `row_parallel_forward`.

One process per GPU, torch.distributed for the plumbing (backend "nccl" over NVLink on B200 boxes; "gloo" in the CPU
tests).  The reference's NCCL backend casts fp32 payloads to bf16 on the wire (model/model_tp_backend.py:119-126);
we reduce in the tensor's own dtype.
"""
from __future__ import annotations
import torch
import torch.distributed as dist

GRANULARITY = 128


def split_ranges(total: int, parts: int, granularity: int = GRANULARITY):
    """Contiguous [first, last) per rank in units of `granularity` channels, as even as possible (earlier ranks get
    the remainder), every rank non-empty when total >= parts * granularity."""
    units = total // granularity
    assert units * granularity == total, f"{total} is not a multiple of {granularity}"
    assert units >= parts, f"cannot split {total} channels over {parts} ranks at granularity {granularity}"
    base, rem = divmod(units, parts)
    out, first = [], 0
    for r in range(parts):
        n = (base + (1 if r < rem else 0)) * granularity
        out.append((first, first + n))
        first += n
    return out


def column_shard(lin, rank: int, world: int, device=None):
    first, last = split_ranges(lin.out_features, world)[rank]
    return lin.tp_slice((True, first, last), device)


def row_shard(lin, rank: int, world: int, device=None):
    first, last = split_ranges(lin.in_features, world)[rank]
    return lin.tp_slice((False, first, last), device)


def all_reduce(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def row_parallel_forward(shard, x_local: torch.Tensor, params: dict, out_dtype=None, group=None) -> torch.Tensor:
    """y = all_reduce( shard(x[:, first:last]) ): each rank applies its own full epilogue to its partial, one sum."""
    y = shard.forward(x_local, params, out_dtype)
    return all_reduce(y, group)


def column_parallel_forward(shard, x: torch.Tensor, params: dict, out_dtype=None) -> torch.Tensor:
    """No communication: every rank keeps its slice of the output channels."""
    return shard.forward(x, params, out_dtype)
