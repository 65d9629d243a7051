"""
Drop-in check at the Python level, run where the reference checkout exists (this container; skipped on the GPU box): the
REFERENCE'S OWN `exllamav3/modules/quant/exl3.py` (LinearEXL3) is loaded from /root/reference -- unmodified, never copied --
with its `ext` import bound to this repo's shim (`exllamav3_b200.ext`), exactly what INTEGRATION.md's three-line patch does.
Every call the reference module then makes into the extension is bound against the shim's real signatures (a wrong argument
count or a missing name fails here), and the reference's dispatch is compared with this repo's own minimal caller (QLinear)
on the same inputs: both must issue the same extension calls with the same argument shapes.  No kernel runs: the shim's entry points are wrapped by recorders after their signatures have been checked.
"""
import importlib.util, inspect, os, sys, types
import pytest
import torch

REF = "/root/reference/exllamav3"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "modules", "quant", "exl3.py")),
                                reason="reference checkout not present (GPU box)")


@pytest.fixture()
def ref_exl3(monkeypatch):
    from exllamav3_b200 import ext
    import refstubs
    calls = []

    def wrap(name):
        real = getattr(ext, name)
        sig = inspect.signature(real)

        def f(*a, **kw):
            sig.bind(*a, **kw)                                    # the reference's argument list fits the shim's signature
            calls.append((name,) + tuple(tuple(t.shape) if isinstance(t, torch.Tensor) else t for t in a))
            return 210
        return f

    for name in ("exl3_gemm", "had_r_128", "reconstruct", "reconstruct_slice", "reconstruct_had_slice", "hgemm"):
        monkeypatch.setattr(ext, name, wrap(name))

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    class Config: pass
    mod("exllamav3")
    mod("exllamav3.model")
    mod("exllamav3.model.config", Config=Config, NullConfig=refstubs.NullConfig)
    mod("exllamav3.ext", exllamav3_ext=ext)                       # <- INTEGRATION.md: the extension module IS the shim
    mod("exllamav3.util", profile_opt=lambda f: f)
    mod("exllamav3.util.tensor", g_tensor_cache=refstubs.g_tensor_cache)
    mod("exllamav3.modules")
    mod("exllamav3.modules.quant")
    mod("exllamav3.modules.quant.exl3_lib")
    mod("exllamav3.modules.quant.exl3_lib.quantize", preapply_had_l=refstubs.preapply_had_l, preapply_had_r=refstubs.preapply_had_r,
        had_k=128, had_n=128)
    spec = importlib.util.spec_from_file_location("exllamav3.modules.quant.exl3", os.path.join(REF, "modules", "quant", "exl3.py"))
    m = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "exllamav3.modules.quant.exl3", m)
    spec.loader.exec_module(m)
    return m, calls


def _tensors(k, n, K=4):
    g = torch.Generator().manual_seed(1)
    tr = torch.randint(0, 32767, (k // 16, n // 16, 16 * K), generator=g, dtype=torch.int32).to(torch.int16)
    return dict(suh=torch.ones(k, dtype=torch.half), svh=torch.ones(n, dtype=torch.half), trellis=tr, mul1=torch.zeros((), dtype=torch.int))


def _qlin(t, **kw):
    from exllamav3_b200 import QLinear
    return QLinear(t["trellis"], t["suh"], t["svh"], mul1=True, **kw)


def test_reference_linear_exl3_runs_on_the_shim_and_matches_the_mirror(ref_exl3):
    ref_mod, calls = ref_exl3
    from exllamav3_b200 import qlinear
    assert ref_mod.AUTO_RECONSTRUCT_THRESHOLD == qlinear.KERNEL_MAX_ROWS == 144
    assert ref_mod.MAX_RECONSTRUCT_SLICE_N == qlinear.DENSE_WINDOW_COLS == 32768
    k, n = 256, 384
    ref_lin = ref_mod.LinearEXL3(None, k, n, key="q", **_tensors(k, n))          # constructs ext.BC_LinearEXL3 from the shim
    mir_lin = _qlin(_tensors(k, n))
    from exllamav3_b200 import ext
    assert isinstance(ref_lin.bc, ext.BC_LinearEXL3) and ref_lin.K == 4 and ref_lin.mul1 and not ref_lin.mcg
    for (shape, params, out_dtype) in (((1, k), {}, None), ((144, k), {}, torch.float), ((3, 7, k), {}, None),
                                       ((145, k), {}, None), ((1, k), {"reconstruct": True}, None), ((1024, k), {}, torch.float)):
        x = torch.zeros(shape, dtype=torch.half)
        calls.clear(); yr = ref_lin.forward(x, params, out_dtype); seq_ref = list(calls)
        calls.clear(); ym = mir_lin.forward(x, params, out_dtype); seq_mir = list(calls)
        assert seq_ref == seq_mir and len(seq_ref) >= 1, (shape, params)            # same extension calls, same argument shapes
        assert yr.shape == ym.shape == tuple(shape[:-1]) + (n,) and yr.dtype == ym.dtype
    # the kernel path hands exl3_gemm the reference's 10 arguments: (A, B, C, suh, A_had, svh, -1, mcg, mul1, 0)
    calls.clear(); ref_lin.forward(torch.zeros((2, k), dtype=torch.half), {})
    assert calls[0][0] == "exl3_gemm" and calls[0][7:] == (-1, False, True, 0)


def test_reference_wide_output_slicing_matches_the_mirror(ref_exl3, monkeypatch):
    ref_mod, calls = ref_exl3
    from exllamav3_b200 import qlinear
    monkeypatch.setattr(ref_mod, "MAX_RECONSTRUCT_SLICE_N", 256)
    monkeypatch.setattr(qlinear, "DENSE_WINDOW_COLS", 256)
    k, n = 128, 640
    ref_lin = ref_mod.LinearEXL3(None, k, n, **_tensors(k, n)); mir_lin = _qlin(_tensors(k, n))
    for rows in (200, 2048):
        x = torch.zeros((rows, k), dtype=torch.half)
        calls.clear(); ref_lin.forward(x, {}); a = list(calls)
        calls.clear(); mir_lin.forward(x, {}); b = list(calls)
        assert a == b and sum(c[0] == "hgemm" for c in a) == 3


def test_reference_multilinear_tables_match_the_mirror():
    """modules/multilinear.py builds the pointer tables exl3_mgemm reads; pointer_tables must build the same ones."""
    src = open(os.path.join(REF, "modules", "multilinear.py")).read()
    ns = {}
    exec(compile(src.replace("from . import Linear", "Linear = object"), "multilinear_ref", "exec"), ns)     # executed in memory, not copied
    from exllamav3_b200 import pointer_tables
    k, n = 128, 256
    inners = [_qlin(_tensors(k, n)) for _ in range(3)]
    for i in inners:
        i.quant_type = "exl3"                                   # the attribute the reference's table builder asserts on

    class Lin:                                                  # the reference wraps LinearEXL3 in modules.Linear (.inner)
        def __init__(self, inner):
            self.inner, self.quant_type, self.softcap, self.post_scale = inner, "exl3", 0.0, 1.0
            self.in_features, self.out_features = inner.in_features, inner.out_features
    ref_ml = ns["MultiLinear"]("cpu", [Lin(i) for i in inners])
    ml = dict(zip(("ptrs_trellis", "ptrs_suh", "ptrs_svh"), pointer_tables("cpu", inners)))
    for a in ("ptrs_trellis", "ptrs_suh", "ptrs_svh"):
        assert torch.equal(getattr(ref_ml, a), ml[a]) and ml[a].dtype == torch.long
    assert (ref_ml.K, ref_ml.mcg, ref_ml.mul1, ref_ml.in_features, ref_ml.out_features) == (4, False, True, k, n)


def test_every_reference_call_site_of_the_qgemm_surface_fits_the_shim():
    """Static audit: every `ext.<op>(...)` call in the reference's Python sources, for the ops of the qgemm path, is bound
    (by argument count and keyword names) against the shim's signature.  SURVEY.md 8b lists these callers."""
    import ast
    from exllamav3_b200 import ext
    ops = ["exl3_gemm", "exl3_mgemm", "reconstruct", "reconstruct_slice", "reconstruct_had_slice", "had_r_128", "hgemm",
           "exl3_gemv", "g_get_cc", "g_get_num_sms", "exl3_gemv_int8_max_k", "exl3_gemm_num_kernel_shapes", "exl3_gemm_shape_compat"]
    sigs = {o: inspect.signature(getattr(ext, o)) for o in ops}
    seen = {o: 0 for o in ops}
    files = 0
    for root, _, names in os.walk(os.path.dirname(REF)):            # the package, science/ (qgemm_benchmark.py), tests/, eval/
        if "exllamav3_ext" in root:
            continue
        for nm in names:
            if not nm.endswith(".py"):
                continue
            path = os.path.join(root, nm)
            try:
                tree = ast.parse(open(path).read())
            except SyntaxError:
                continue
            files += 1
            for node in ast.walk(tree):
                if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in sigs
                        and isinstance(node.func.value, ast.Name) and node.func.value.id == "ext"):
                    if any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords):
                        continue
                    op = node.func.attr
                    try:
                        sigs[op].bind(*([None] * len(node.args)), **{k.arg: None for k in node.keywords})
                    except TypeError as e:
                        raise AssertionError(f"{path}:{node.lineno}: ext.{op} call does not fit the shim: {e}")
                    seen[op] += 1
    assert files > 50
    assert seen["exl3_gemm"] >= 1 and seen["exl3_gemm_num_kernel_shapes"] >= 1 and seen["exl3_gemm_shape_compat"] >= 1   # science/qgemm_benchmark.py
    # the call sites SURVEY.md 8b names exist and were checked
    assert seen["exl3_mgemm"] >= 8 and seen["hgemm"] >= 3 and seen["had_r_128"] >= 2 and seen["reconstruct"] >= 2
    assert seen["reconstruct_had_slice"] >= 2 and seen["reconstruct_slice"] >= 1 and seen["exl3_gemv_int8_max_k"] >= 1


def test_reference_use_mgemm_policy_with_the_shims_answer(monkeypatch):
    """model/config.py:48-64, the reference's OWN policy function, evaluated with this shim's exl3_gemv_int8_max_k (0): k+v and
    gate+up stay fused for every bitrate -- the launch list bench.py times (161 launches) -- whereas with the reference
    extension's answer on Blackwell (6) the wide 4-bpw gate+up pair is unfused (225-launch variant, bench.py --no-fuse)."""
    from exllamav3_b200 import ext

    def mod(name, **attrs):
        m = types.ModuleType(name); m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        monkeypatch.setitem(sys.modules, name, m)

    mod("exllamav3"); mod("exllamav3.model"); mod("exllamav3.util")
    mod("exllamav3.util.rope", RopeSettings=object, RopeStyle=object)
    mod("exllamav3.loader", SafetensorsCollection=object)
    mod("exllamav3.util.file", read_dict=None, no_value=None, no_default=None)
    mod("exllamav3.ext", exllamav3_ext=ext)
    for v in ("EXL3_INT8_GEMV", "EXL3_MGEMM_K_THRESHOLD", "EXL3_MGEMM_N_THRESHOLD"):
        monkeypatch.delenv(v, raising=False)
    spec = importlib.util.spec_from_file_location("exllamav3.model.config", os.path.join(REF, "model", "config.py"))
    cfg = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "exllamav3.model.config", cfg)
    spec.loader.exec_module(cfg)
    ip = cfg.InferParams()
    monkeypatch.setattr(ext, "_check", lambda rc: 148)                      # no device here: the index check is not under test
    assert ext.exl3_gemv_int8_max_k(0) == 0
    for K in range(1, 9):
        assert ip.use_mgemm(K, 14336, mul1=True, device="cuda:0")          # gate/up: fused at every bitrate
        assert ip.use_mgemm(K, 1024, mul1=True, device="cuda:0")           # k/v
        assert ip.use_mgemm(K, 14336, mul1=False, device="cuda:0")         # other codebooks are always fused
    monkeypatch.setattr(ext, "exl3_gemv_int8_max_k", lambda dev: 6)        # what the reference's own extension reports on sm_100
    assert not ip.use_mgemm(4, 14336, mul1=True, device="cuda:0")          # wide 4 bpw pair: unfused there
    assert ip.use_mgemm(4, 1024, mul1=True, device="cuda:0")               # narrow outputs stay fused either way
    assert ip.use_mgemm(7, 14336, mul1=True, device="cuda:0")


def test_reference_tp_import_split_equals_tp_slice(ref_exl3):
    """The reference's OWN shard construction (LinearEXL3.tp_export / tp_import_split, modules/quant/exl3.py:268-330), run with a
    trivial in-process stand-in for its SHM producer / consumer transport, against this repo's tp.tp_slice."""
    ref_mod, _ = ref_exl3
    from exllamav3_b200 import tp

    class Producer:
        def send(self, t): return t

    class Consumer:
        def recv(self, t, cuda=False, slice_dim=None, first=None, last=None):
            if t is None:
                return None
            return t if slice_dim is None else t.narrow(slice_dim, first, last - first).contiguous()

    k, n = 512, 768
    tens = _tensors(k, n)
    g = torch.Generator().manual_seed(5)
    tens["suh"] = torch.randn(k, generator=g).half(); tens["svh"] = torch.randn(n, generator=g).half()
    bias = torch.randn(n, generator=g).half()
    ref_full = ref_mod.LinearEXL3(None, k, n, bias=bias, out_dtype=torch.float, **tens)
    mir_full = _qlin(tens, bias=bias, out_dtype=torch.float)
    exported = ref_full.tp_export(None, Producer())
    ctx = {"consumer": Consumer(), "device": "cpu"}
    world = 3
    splits = [(True, a, b) for (a, b) in tp.split_ranges(n, world)] + [(False, a, b) for (a, b) in tp.split_ranges(k, world)] + [None]
    for split in splits:
        r = ref_mod.LinearEXL3.tp_import_split(ctx, exported, None, split)
        m = tp.tp_slice(mir_full, split)
        assert (r.in_features, r.out_features, r.K, r.mcg, r.mul1, r.out_dtype) == (m.in_features, m.out_features, m.K, m.mcg, m.mul1, m.out_dtype)
        assert torch.equal(r.trellis, m.trellis) and torch.equal(r.suh, m.suh) and torch.equal(r.svh, m.svh)
        assert (r.bias is None) == (m.bias is None) and (r.bias is None or torch.equal(r.bias, m.bias))
        assert m.trellis.is_contiguous()
