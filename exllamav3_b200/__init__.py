"""
exllamav3_b200 -- Blackwell-native (sm_100a) EXL3 quantized-GEMM path.

    ext           reference-named operator surface (exl3_gemm, exl3_mgemm, reconstruct*, had_r_128, hgemm,
                  BC_LinearEXL3) over the C ABI of libexl3b200.so
    LinearEXL3    mirror of exllamav3.modules.quant.exl3.LinearEXL3
    MultiLinear   pointer tables for exl3_mgemm (exllamav3.modules.multilinear.MultiLinear)
    tp            column/row tensor-parallel shard + NCCL all-reduce around the row-parallel output
"""
import sys as _sys

# `python -m exllamav3_b200.build` imports this package before the library exists: only then skip the operator surface
# (everything else fails loudly in ext.py when libexl3b200.so is missing -- there is no fallback implementation)
if "exllamav3_b200.build" not in getattr(_sys, "orig_argv", ()):
    from . import ext
    from .linear_exl3 import LinearEXL3, AUTO_RECONSTRUCT_THRESHOLD
    from .multilinear import MultiLinear

__all__ = ["ext", "LinearEXL3", "MultiLinear", "AUTO_RECONSTRUCT_THRESHOLD"]
