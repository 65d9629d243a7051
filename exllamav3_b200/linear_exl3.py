"""
LinearEXL3 -- host-side mirror of the reference's `exllamav3.modules.quant.exl3.LinearEXL3`
(modules/quant/exl3.py:14-389) running on exllamav3_b200.ext instead of exllamav3_ext.

Same constructor arguments, attributes and methods for the qgemm path: forward (kernel path for
rows <= AUTO_RECONSTRUCT_THRESHOLD, reconstruct + hgemm above), reconstruct_hgemm, get_inner_weight_tensor,
get_weight_tensor, unpack_bf, and the tensor-parallel column/row slicing of tp_import_split
(modules/quant/exl3.py:284-330), exposed as `tp_slice` because the reference's SHM producer/consumer transport
(model/model_tp_shared.py) is out of scope.
"""
from __future__ import annotations
import os
import torch
from . import ext

AUTO_RECONSTRUCT_THRESHOLD = 144          # modules/quant/exl3.py:10
MAX_RECONSTRUCT_SLICE_N = 32768           # modules/quant/exl3.py:11
RECONSTRUCT_SLICE_GRANULARITY_N = 128     # modules/quant/exl3.py:12

no_fused_reconstruct = os.environ.get("EXL3_NO_FUSED_RECONSTRUCT", "0") != "0"   # doc/env_vars.md:156


class InferParams:
    """Subset of model/config.py:11-64 read by the qgemm path."""
    def __init__(self, no_reconstruct: bool = False):
        self.no_reconstruct = no_reconstruct


class NullConfig:
    def __init__(self):
        self.infer_params = InferParams()


class _TensorCache:
    """Process-global scratch cache keyed by (device, shape, dtype), as util/tensor.py:213-241 (g_tensor_cache)."""
    def __init__(self):
        self._d = {}

    def get(self, device, shape, dtype):
        key = (str(device), tuple(shape), dtype)
        if key not in self._d:
            self._d[key] = torch.empty(shape, dtype=torch.half, device=device)
        return self._d[key]


g_tensor_cache = _TensorCache()


class LinearEXL3:

    quant_type: str = "exl3"

    def __init__(
        self,
        config,
        in_features: int,
        out_features: int,
        scale: torch.Tensor | None = None,
        su: torch.Tensor | None = None,
        sv: torch.Tensor | None = None,
        suh: torch.Tensor | None = None,
        svh: torch.Tensor | None = None,
        trellis: torch.Tensor | None = None,
        mcg: torch.Tensor | None = None,
        mul1: torch.Tensor | None = None,
        bias: torch.Tensor | None = None,
        out_dtype: torch.dtype | None = None,
        transformers_fix: bool = False,
        key: str | None = None,
    ):
        assert scale is None, "scale is no longer used"
        assert su is not None or suh is not None, "either su (packed) or suh (unpacked) is required"
        assert sv is not None or svh is not None, "either sv (packed) or svh (unpacked) is required"
        assert trellis is not None, "trellis is required"
        if su is not None: assert su.dtype == torch.int16, "su is wrong datatype"
        if sv is not None: assert sv.dtype == torch.int16, "sv is wrong datatype"
        if suh is not None: assert suh.dtype == torch.half, "suh is wrong datatype"
        if svh is not None: assert svh.dtype == torch.half, "svh is wrong datatype"
        assert trellis.dtype == torch.int16, "trellis is wrong datatype"
        assert len(trellis.shape) == 3, "trellis must have dim = 3"

        if bias is not None and bias.dtype == torch.float:
            bias = bias.to(torch.half)

        self.config = config if config is not None else NullConfig()
        self.transformers_fix = transformers_fix
        self.key = key
        self.su = None
        self.sv = None
        self.suh = suh if suh is not None else self.unpack_bf(su)
        self.svh = svh if svh is not None else self.unpack_bf(sv)
        self.trellis = trellis
        self.K = trellis.shape[-1] // 16
        self.in_features = in_features
        self.out_features = out_features
        self.bias = bias
        self.out_dtype = out_dtype
        self.default_out_dtype = out_dtype or torch.half
        self.mcg_tensor = mcg
        self.mul1_tensor = mul1
        self.mcg = mcg is not None
        self.mul1 = mul1 is not None
        self._fused_reconstruct = None
        self.bsz1_xh_args = (self.trellis.device, (1, self.in_features), self.out_dtype)
        self.bc = ext.BC_LinearEXL3(
            self.trellis, self.suh, self.svh, self.K, self.bias, self.mcg, self.mul1,
            g_tensor_cache.get(*self.bsz1_xh_args))

    def get_tensors(self, key: str):
        return {
            f"{key}.{sub}": t.contiguous()
            for sub, t in [("su", self.su), ("sv", self.sv), ("suh", self.suh), ("svh", self.svh),
                           ("trellis", self.trellis), ("bias", self.bias),
                           ("mcg", self.mcg_tensor), ("mul1", self.mul1_tensor)] if t is not None
        }

    def forward(self, x: torch.Tensor, params: dict, out_dtype: torch.dtype | None = None) -> torch.Tensor:
        # modules/quant/exl3.py:114-139
        if "ovr" in params:                      # per-call module override table (exl3.py:121-124)
            ovr = params["ovr"]
            if self.key in ovr and getattr(ovr[self.key], "inner", None) is not self:
                return ovr[self.key].forward(x, params, out_dtype)
        assert x.is_contiguous(), f"LinearEXL3 {self.key}: non-contiguous input {tuple(x.shape)}"
        reconstruct = params.get("reconstruct")
        if not reconstruct:
            rows = x.numel() // x.shape[-1]
            if rows <= AUTO_RECONSTRUCT_THRESHOLD or self.config.infer_params.no_reconstruct:
                dtype = out_dtype or self.default_out_dtype
                return self.bc.run_alloc(x, self.out_features, dtype == torch.float)
        return self.reconstruct_hgemm(x, out_dtype)

    def unpack_bf(self, bitfield: torch.Tensor):
        # modules/quant/exl3.py:142-158 (legacy packed sign bitfields -> +-1 fp16)
        device = bitfield.device
        bitfield = bitfield.view(torch.uint16).to(torch.int)
        masks = (1 << torch.arange(16)).to(bitfield.device)
        expanded = ((bitfield.unsqueeze(-1) & masks) > 0).flatten()
        expanded = 1.0 - expanded.to(torch.float16) * 2.0
        return expanded.contiguous().to(device)

    def reconstruct_hgemm(self, x: torch.Tensor, out_dtype):
        # modules/quant/exl3.py:161-218
        shape = x.shape
        rows = x.numel() // shape[-1]
        out_shape = shape[:-1] + (self.out_features,)
        x = x.view(rows, self.in_features)
        y = torch.empty(out_shape, dtype=out_dtype or self.default_out_dtype, device=x.device)
        y_ = y.view(rows, self.out_features)

        if self._fused_reconstruct is None:
            self._fused_reconstruct = (
                self.in_features % 128 == 0 and self.out_features % 128 == 0 and not no_fused_reconstruct)
        use_fused = self._fused_reconstruct and rows >= 1024

        if use_fused:
            xh = x
        else:
            xh = torch.empty_like(x)
            ext.had_r_128(x, xh, self.suh, None, 1.0)

        if self.out_features <= MAX_RECONSTRUCT_SLICE_N:
            w = torch.empty((self.in_features, self.out_features), dtype=torch.half, device=self.trellis.device)
            if use_fused:
                ext.reconstruct_had_slice(w, self.trellis, self.suh, self.svh, self.K, self.mcg, self.mul1, 0)
            else:
                ext.reconstruct(w, self.trellis, self.K, self.mcg, self.mul1)
            ext.hgemm(xh, w, y_)
        else:
            numel_ = self.in_features * MAX_RECONSTRUCT_SLICE_N
            w_ = torch.empty((numel_,), dtype=torch.half, device=self.trellis.device)
            for n_start in range(0, self.out_features, MAX_RECONSTRUCT_SLICE_N):
                n_end = min(n_start + MAX_RECONSTRUCT_SLICE_N, self.out_features)
                numel = self.in_features * (n_end - n_start)
                w = w_[:numel].view(self.in_features, n_end - n_start)
                if use_fused:
                    ext.reconstruct_had_slice(
                        w, self.trellis, self.suh, self.svh[n_start:], self.K, self.mcg, self.mul1, n_start)
                else:
                    ext.reconstruct_slice(w, self.trellis, self.K, self.mcg, self.mul1, n_start)
                ext.hgemm(xh, w, y_[:, n_start:n_end])

        if not use_fused:
            ext.had_r_128(y_, y_, None, self.svh, 1.0)
        if self.bias is not None:
            y += self.bias
        return y

    def get_inner_weight_tensor(self):
        w = torch.empty((self.in_features, self.out_features), dtype=torch.half, device=self.trellis.device)
        ext.reconstruct(w, self.trellis, self.K, self.mcg, self.mul1)
        return w

    def get_weight_tensor(self):
        # modules/quant/exl3.py:227-237; Hadamards as fp32 matmuls like quantize.py:340-357
        from .hadamard import preapply_had_l, preapply_had_r
        w = self.get_inner_weight_tensor()
        w = preapply_had_l(w, 128)
        w *= self.suh.unsqueeze(1)
        w = preapply_had_r(w, 128)
        w *= self.svh.unsqueeze(0)
        return w

    def get_bias_tensor(self):
        return self.bias

    # ---- tensor parallel ---------------------------------------------------------------------------------------

    def tp_slice(self, split, device=None) -> "LinearEXL3":
        """
        Column (split_out=True) or row (split_out=False) shard [first, last) in channels, multiples of 128.
        Same slicing as LinearEXL3.tp_import_split (modules/quant/exl3.py:284-330): column split slices
        trellis[:, first/16:last/16], svh, bias and keeps suh; row split slices trellis[first/16:last/16], suh,
        keeps svh and gives the bias only to the shard with first == 0.
        """
        split_out, first, last = split if split is not None else (True, 0, self.out_features)
        assert first % 128 == 0 and last % 128 == 0, "TP split granularity is 128 channels (modules/linear.py:645-656)"
        dev = device or self.trellis.device
        mv = lambda t: None if t is None else t.to(dev).contiguous()
        if split_out:
            suh, svh = mv(self.suh), mv(self.svh[first:last])
            trellis = mv(self.trellis[:, first // 16: last // 16, :])
            bias = mv(self.bias[first:last]) if self.bias is not None else None
            in_f, out_f = self.in_features, last - first
        else:
            suh, svh = mv(self.suh[first:last]), mv(self.svh)
            trellis = mv(self.trellis[first // 16: last // 16])
            bias = mv(self.bias) if (self.bias is not None and first == 0) else None
            in_f, out_f = last - first, self.out_features
        return LinearEXL3(None, in_f, out_f, suh=suh, svh=svh, trellis=trellis, mcg=mv(self.mcg_tensor),
                          mul1=mv(self.mul1_tensor), bias=bias, out_dtype=self.out_dtype, key=self.key)
