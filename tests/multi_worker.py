"""Worker of tests/test_gpu_multi.py (one process per GPU, launched with torch.distributed.run): a wide A16W4 layer is
built identically on every rank, sharded by output column, and (1) the NCCL path -- ColumnShardedLinear.forward = local
kernel + all-gather -- and (2) the fused path -- forward_scatter: the decode kernel's epilogue stores into every rank's
symmetric output buffer, one barrier -- must both reproduce the unsharded layer's output bit for bit on every rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gemlite_b200 import _cabi  # noqa: E402
from gemlite_b200.helper import A16W4_HQQ_INT, A8W8_INT8_dynamic  # noqa: E402
from gemlite_b200.sharding import ColumnShardedLinear, SymmetricOutputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    try:
        K, N = 4096, 4096 * world
        g = torch.Generator(device=dev).manual_seed(0)                       # same seed => same global layer on every rank
        W_q = torch.randint(0, 16, (N, K), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
        s = (torch.rand(N * K // 128, 1, generator=g, device=dev) * 0.01 + 0.001).half()
        z = (torch.rand(N * K // 128, 1, generator=g, device=dev) * 15).half()
        full = A16W4_HQQ_INT(device=dev, dtype=torch.float16).from_weights(W_q, s, z)
        shard = ColumnShardedLinear(full, rank, world)
        sym = SymmetricOutputs(3, N, torch.float16, dev)
        for M in (1, 5, 200):
            x = (torch.randn(M, K, generator=g, device=dev) / 10).half()
            want = full(x)
            got = shard(x)                                                    # local kernel + NCCL all-gather
            assert torch.equal(got, want), ("nccl", M, rank)
        for rep in range(3):                                                  # fused: M == 1 only
            x = (torch.randn(1, K, generator=g, device=dev) / 10).half()
            want = full(x)
            sym.barrier()
            shard.forward_scatter(x, sym, row=rep)
            assert _cabi.last_kernel() == "gemv_w4_mma"
            sym.barrier()
            torch.cuda.synchronize()
            assert torch.equal(sym.buf[rep].view(1, N), want), ("fused", rep, rank)
        # a layer kind without a multi-destination kernel must say so (the caller then uses forward())
        W = (torch.randn(512, 1024, generator=g, device=dev) / 10).half()
        a8 = ColumnShardedLinear(A8W8_INT8_dynamic(device=dev, dtype=torch.float16).from_weights(W), rank, world)
        try:
            a8.forward_scatter(torch.zeros(1, 1024, dtype=torch.float16, device=dev), SymmetricOutputs(1, 512, torch.float16, dev))
            raise AssertionError("expected NotImplementedError")
        except (NotImplementedError, TypeError):
            pass
        dist.barrier()
        if rank == 0:
            print("MULTI_OK", world)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
