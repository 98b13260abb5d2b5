"""Row (e) host logic on CPU: column-shard a packed layer over world_size 2 (gloo), run every shard, all-gather the
outputs, and require the result to be IDENTICAL to the unsharded layer.  The per-shard compute in this CPU test is
the oracle (the product has no CPU path); what is under test is `gemlite_b200.sharding` -- the slicing of packed
weights / scales / zeros / bias (incl. the MX and per-channel layouts and the reference's placeholders) and the
order of the gathered column blocks."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gemlite_b200 import sharding as S
from oracle import gemlite_oracle as O
from oracle import recipes as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _cases():
    g = torch.Generator().manual_seed(0)
    N, K = 128, 256
    out = {}
    W_q = torch.randint(0, 16, (N, K), generator=g, dtype=torch.int32).to(torch.uint8)
    scales = (torch.rand(N * K // 64, 1, generator=g) * 0.01 + 0.001).half()
    zeros = (torch.rand(N * K // 64, 1, generator=g) * 15).half()
    bias = (torch.randn(N, generator=g) / 10).half()
    out["a16w4_gs64_fma_bias"] = O.prepare(W_q, scales, zeros, W_nbits=4, group_size=64, in_features=K, out_features=N, bias=bias)
    out["a16w4_gs64_sym"] = O.prepare(W_q, scales, None, W_nbits=4, group_size=64, in_features=K, out_features=N)
    W = torch.randn(N, K, generator=g) / 10
    out["a16w8_int8"] = R.a16w8(W.half(), torch.float16)
    out["a8w8_int8_dynamic"] = R.a8w8_dynamic(W.half(), torch.float16)
    return out


def _shard_packed(p, rank, world):
    q = copy.copy(p)
    N = p.out_features
    q.W_q, q.scales, q.zeros, q.bias = S.shard_tensors(p.W_q, p.scales, p.zeros, p.bias, N, rank, world,
                                                      contiguous=bool(p.data_contiguous))
    q.out_features = N // world
    return q


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = {}
        for name, p in _cases().items():
            x = (torch.randn(5, p.in_features, generator=torch.Generator().manual_seed(3)) / 10).half()
            sx = None
            if p.scaled_activations:
                x, sx = O.quantize_activations_per_token(x, p.W_q.dtype)
            full = O.forward(p, x, sx)
            local = O.forward(_shard_packed(p, rank, world), x, sx)
            assert local.shape == (5, p.out_features // world)
            gathered = S.gather_columns(local)
            ok[name] = bool(torch.equal(gathered, full))
            # 3-d activations keep their leading dims through the gather
            g3 = S.gather_columns(local.view(1, 5, -1))
            ok[name + "/3d"] = tuple(g3.shape) == (1, 5, p.out_features) and bool(torch.equal(g3[0], full))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


def test_column_shards_gather_to_the_unsharded_result_gloo_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert len(results) == world
        for rank in range(world):
            assert all(results[rank].values()), (rank, dict(results[rank]))


def test_column_range_and_placeholders():
    assert S.column_range(4096, 1, 4) == (1024, 2048)
    with pytest.raises(NotImplementedError):
        S.column_range(4096 + 32, 0, 4)
    with pytest.raises(NotImplementedError):
        S.column_range(96, 0, 2)          # 48 columns per shard: not a whole number of 32-column slabs
    # the reference's placeholders (core.py:470-475) travel through unchanged
    W = torch.zeros(8, 64, dtype=torch.int32)
    empty = torch.tensor([[]], dtype=torch.int32)
    zscalar = torch.tensor(7, dtype=torch.int32)
    Wl, sl, zl, bl = S.shard_tensors(W, empty, zscalar, None, 64, 1, 2)
    assert Wl.shape == (8, 32) and Wl.is_contiguous() and sl is empty and zl is zscalar and bl is None
    # MX scales are stored (N, K/gs) with strides (1, N): the shard keeps that orientation
    s_mx = torch.arange(4 * 64, dtype=torch.uint8).view(4, 64).t()
    _, sl, _, _ = S.shard_tensors(W, s_mx, empty, None, 64, 1, 2)
    assert sl.shape == (32, 4) and sl.stride() == (1, 32) and torch.equal(sl, s_mx[32:])
    # K/gs == N (N = 64, K = 2048, gs = 32): the shape alone cannot tell the group axis from the feature axis --
    # the layer kind (mx flag) or, without it, the storage order does (sliced along K before)
    s_sq = torch.arange(64 * 64, dtype=torch.int32).to(torch.uint8).view(64, 64).t()      # MX view: (N, K/gs), strides (1, N)
    for flag in (True, None):
        _, sl, _, _ = S.shard_tensors(W, s_sq, empty, None, 64, 1, 2, mx=flag)
        assert sl.shape == (32, 64) and sl.stride() == (1, 32) and torch.equal(sl, s_sq[32:]), flag
    s_grp = torch.arange(64 * 64, dtype=torch.float32).view(64, 64)                       # ordinary (K/gs, N), contiguous
    for flag in (False, None):
        _, sl, _, _ = S.shard_tensors(W, s_grp, empty, None, 64, 1, 2, mx=flag)
        assert sl.shape == (64, 32) and sl.is_contiguous() and torch.equal(sl, s_grp[:, 32:]), flag


def test_column_sharded_linear_builds_a_consistent_local_layer():
    """ColumnShardedLinear on a CPU-resident packed layer (built through load_state_dict from the oracle's packed
    tensors -- pack() itself needs the GPU): the local module carries the shard's tensors, the parent's metadata and
    out_features / world columns.  (Its forward is the ordinary C-ABI call and is covered by the GPU suite.)"""
    from gemlite_b200 import GemLiteLinear
    p = _cases()["a16w4_gs64_fma_bias"]
    layer = GemLiteLinear.__new__(GemLiteLinear)
    torch.nn.Module.__init__(layer)
    layer.load_state_dict({"W_q": p.W_q, "bias": p.bias, "scales": p.scales, "zeros": p.zeros,
                           "metadata": torch.tensor(p.meta_args(), dtype=torch.int32),
                           "orig_shape": torch.tensor([p.out_features, p.in_features], dtype=torch.int32)})
    world = 2
    for rank in range(world):
        sh = S.ColumnShardedLinear(layer, rank, world)
        b, e = S.column_range(p.out_features, rank, world)
        assert (sh.in_features, sh.out_features) == (p.in_features, p.out_features)
        assert (sh.local.in_features, sh.local.out_features) == (p.in_features, p.out_features // world)
        assert sh.local.get_meta_args() == layer.get_meta_args()
        assert torch.equal(sh.local.W_q, p.W_q[:, b:e]) and sh.local.W_q.is_contiguous()
        assert torch.equal(sh.local.scales, p.scales[:, b:e]) and torch.equal(sh.local.zeros, p.zeros[:, b:e])
        assert torch.equal(sh.local.bias, p.bias[b:e])
