"""
Test harness only: the few names the REFERENCE'S OWN modules import from the rest of its package (config defaults, the
process-global scratch cache, the Hadamard helpers of its quantizer), provided as minimal stand-ins so that
/root/reference/exllamav3/modules/quant/exl3.py can be loaded unmodified with its `ext` bound to exllamav3_b200.ext
(tests/test_reference_callsites.py).  Nothing in the product imports this file.
"""
import numpy as np
import torch


class InferParams:
    no_reconstruct = False


class NullConfig:
    def __init__(self):
        self.infer_params = InferParams()


class TensorCache:
    def __init__(self):
        self.d = {}

    def get(self, device, shape, dtype):
        return self.d.setdefault((str(device), tuple(shape), dtype), torch.empty(shape, dtype=torch.half, device=device))


g_tensor_cache = TensorCache()


def _h(n, ref):
    from oracle import exl3_oracle as orc
    assert n == 128
    return torch.from_numpy(orc.hadamard_matrix_128() / np.sqrt(128.0)).to(ref.device, ref.dtype)


def preapply_had_l(x, had_dim):
    k, n = x.shape
    return (_h(had_dim, x) @ x.view(k // had_dim, had_dim, n)).view(k, n)


def preapply_had_r(x, had_dim):
    k, n = x.shape
    return (x.view(k, n // had_dim, had_dim) @ _h(had_dim, x)).view(k, n)
