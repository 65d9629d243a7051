"""
N > 1 host logic on CPU: world_size-2 `gloo` processes build column / row shards with the reference's slicing rules
and run the collective plumbing of exllamav3_b200.tp.  There is no GPU here, so each rank's local matmul is computed
by the oracle (checker) on that rank's SHARD tensors -- what is under test is the sharding + all-reduce logic:
column shards concatenate to, and row shards sum to, the unsharded result.
"""
import os, sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from exllamav3_b200 import QLinear, tp
        from oracle import exl3_oracle as orc
        k, n, K, cb, m = 512, 768, 4, 2, 3
        tr, suh, svh, x = orc.make_synthetic(k, n, K, m=m)
        lin = QLinear(torch.from_numpy(tr), torch.from_numpy(suh), torch.from_numpy(svh), mul1=True)
        full = orc.exl3_gemm_f64(x, tr, suh, svh, K, cb)

        def local_gemm(shard, x_local):       # oracle stands in for the CUDA kernel on this GPU-less box
            y = orc.exl3_gemm_f64(x_local, shard.trellis.numpy(), shard.suh.numpy(), shard.svh.numpy(), shard.K, cb)
            return torch.from_numpy(y)

        # column parallel: no communication, gather only to verify
        cs = tp.column_shard(lin, rank, world)
        yc = local_gemm(cs, x)
        err_c = float((tp.gather_columns(yc, n) - torch.from_numpy(full)).abs().max())

        # row parallel: partial with full epilogue on every rank, one all-reduce
        rs = tp.row_shard(lin, rank, world)
        first, last = tp.split_ranges(k, world)[rank]
        yr = tp.all_reduce(local_gemm(rs, np.ascontiguousarray(x[:, first:last])))
        err_r = float((yr - torch.from_numpy(full)).abs().max())
        ret[rank] = (err_c, err_r, cs.out_features, rs.in_features)
    finally:
        dist.destroy_process_group()


def test_tp_column_and_row_sharding_world2():
    world, port = 2, 29517 + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err_c, err_r, out_f, in_f = ret[rank]
        assert out_f == 384 and in_f == 256
        assert err_c < 1e-9, err_c            # identical math, just partitioned along n
        assert err_r < 2e-3, err_r            # fp16 rounding of xh is per 128-block => partial sums differ only in fp64 order


def _worker_rank_ordered(rank, world, port, ret):
    """The fused row-parallel epilogue's arithmetic (gemm_tc_i8_body.cuh emit_rows, AR = true), restated on the host: every
    rank holds every rank's finished fp32 partial and adds them in RANK order, so all ranks end with bit-identical sums."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from exllamav3_b200 import QLinear, tp
        from oracle import exl3_oracle as orc
        k, n, K, cb, m = 768, 512, 4, 2, 2
        tr, suh, svh, x = orc.make_synthetic(k, n, K, m=m)
        lin = QLinear(torch.from_numpy(tr), torch.from_numpy(suh), torch.from_numpy(svh), mul1=True)
        rs = tp.row_shard(lin, rank, world)
        first, last = tp.split_ranges(k, world)[rank]
        part = torch.from_numpy(orc.exl3_gemm_f64(np.ascontiguousarray(x[:, first:last]), rs.trellis.numpy(), rs.suh.numpy(),
                                                  rs.svh.numpy(), rs.K, cb)).float()           # the rank's fp32 partial
        slots = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(slots, part)                                   # = every peer storing into slot [its rank]
        acc = torch.zeros_like(part)
        for j in range(world):
            acc = acc + (part if j == rank else slots[j])             # rank order, own partial from registers
        ref = torch.from_numpy(orc.exl3_gemm_f64(x, tr, suh, svh, K, cb))
        everyone = [torch.empty_like(acc) for _ in range(world)]
        dist.all_gather(everyone, acc)
        # column-parallel with UNEVEN shards (4 units of 128 over 3 ranks: 256 + 128 + 128) gathered to full width
        cs = tp.column_shard(lin, rank, world)
        yc = torch.from_numpy(orc.exl3_gemm_f64(x, cs.trellis.numpy(), cs.suh.numpy(), cs.svh.numpy(), cs.K, cb))
        err_cols = float((tp.gather_columns(yc, n) - ref).abs().max())
        ret[rank] = (float((acc.double() - ref).abs().max() / ref.abs().max()),
                     all(torch.equal(everyone[0], e) for e in everyone) and err_cols < 1e-9 and cs.out_features == (256 if rank == 0 else 128))
    finally:
        dist.destroy_process_group()


def test_tp_rank_ordered_sum_is_identical_on_all_ranks_world3():
    world, port = 3, 29917 + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_rank_ordered, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, identical = ret[rank]
        assert err < 2e-3 and identical


def _worker_fused_setup(rank, world, port, ret):
    """tp.enable_fused_allreduce's collective plumbing with the CUDA IPC calls stubbed: every rank must hand tp_attach the
    world handles concatenated in RANK order, and nobody may proceed before everybody has attached (barrier)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from exllamav3_b200 import ext, tp
        seen = {}
        ext.tp_alloc = lambda r, w, me: seen.update(alloc=(r, w, me)) or bytes([0x40 + r]) * ext.TP_HANDLE_BYTES
        ext.tp_attach = lambda handles, w: seen.update(attach=(bytes(handles), w))
        ext.tp_free = lambda: seen.update(freed=True)
        tp.enable_fused_allreduce(max_elems=2048)
        want = b"".join(bytes([0x40 + j]) * ext.TP_HANDLE_BYTES for j in range(world))
        ok = seen["alloc"] == (rank, world, 2048) and seen["attach"] == (want, world) and tp._fused == {"on": True, "world": world, "max_elems": 2048}
        tp.disable_fused_allreduce()
        ok = ok and seen.get("freed") and tp._fused["on"] is False
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_fused_allreduce_setup_plumbing_world3():
    world, port = 3, 30317 + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_fused_setup, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret[r] for r in range(world))


def _worker_expert_parallel(rank, world, port, ret):
    """Expert-parallel MoE down-projection (BASELINE config 5; the reference partitions experts by contiguous index range and
    all-reduces the outputs, modules/block_sparse_mlp.py:1679-1705, no all-to-all): every rank runs exl3_mgemm semantics with
    the expert-range filter [lo, hi) and its LOCAL pointer tables, the weighted partial sums add up to the unsharded result."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import exl3_oracle as orc
        E, k, n, K, m, topk = 4, 256, 128, 4, 1, 3
        mats = [orc.make_synthetic(k, n, K, seed=60 + e, m=m) for e in range(E)]
        trs = [t[0] for t in mats]; suhs = [t[1] for t in mats]; svhs = [t[2] for t in mats]
        rng = np.random.default_rng(3)
        xs = rng.standard_normal((topk, m, k)).astype(np.float16)
        sel = [3, 0, 2]; wts = rng.uniform(0.1, 1.0, topk).astype(np.float16)
        full = orc.exl3_mgemm(xs, trs, suhs, svhs, K, 2, np.float32, indices=sel, weights=wts, bszm_out=topk)[0]
        per = E // world
        lo, hi = rank * per, (rank + 1) * per
        # compaction keeps the slot ORDER but slot j of the shard reads input A[j] (exl3_gemm_kernel.cuh:139-146): the caller
        # passes the inputs of the retained slots, as the reference's expert-parallel path does
        keep = [j for j, q in enumerate(sel) if lo <= q < hi]
        xs_local = np.ascontiguousarray(np.concatenate([xs[keep], np.zeros((topk - len(keep), m, k), dtype=np.float16)], axis=0))
        part = orc.exl3_mgemm(xs_local, trs[lo:hi], suhs[lo:hi], svhs[lo:hi], K, 2, np.float32, indices=sel, weights=wts,
                              min_index=lo, max_index=hi, bszm_out=topk)[0]
        t = torch.from_numpy(np.ascontiguousarray(part)).float()
        dist.all_reduce(t)
        ret[rank] = (float(np.abs(t.numpy() - full).max() / np.abs(full).max()), len(keep))
    finally:
        dist.destroy_process_group()


def test_expert_parallel_partials_sum_to_the_unsharded_result_world2():
    world, port = 2, 30717 + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_expert_parallel, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    assert sorted(ret[r][1] for r in range(world)) == [1, 2]          # experts {0} on rank 0, {3, 2} on rank 1
    for r in range(world):
        assert ret[r][0] < 1e-5, ret[r]


def test_split_ranges():
    from exllamav3_b200 import tp
    assert tp.split_ranges(4096, 4) == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)]
    assert tp.split_ranges(14336, 8)[0] == (0, 1792) and tp.split_ranges(14336, 8)[-1][1] == 14336
    r = tp.split_ranges(1024, 3)          # uneven: 8 units -> 3, 3, 2
    assert [b - a for a, b in r] == [384, 384, 256]
    with pytest.raises(AssertionError):
        tp.split_ranges(256, 4)
