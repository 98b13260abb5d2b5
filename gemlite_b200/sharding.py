"""Column sharding of a packed layer across the GPUs of one box.

The path shards by OUTPUT COLUMN with no data dependency between shards: `y[:, n]` depends only on column
`n` of the packed weights and of its scales / zeros.  Rank r of g owns columns [r*N/g, (r+1)*N/g):
`W_q[:, slice]` (packed `(K/eps, N/g)`), `scales / zeros[:, slice]` (MX: `scales[slice, :]`), `bias[slice]`;
x is replicated.  The only exchange step is ONE all-gather of the local outputs along the last dimension
(NCCL on GPUs; the host logic is backend-agnostic and is covered on CPU with gloo, world_size 2).

Nothing here computes: the shard is an ordinary `GemLiteLinear` over the sliced tensors, its forward is
the same C-ABI call as the unsharded layer.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor


def column_range(n_features: int, rank: int, world: int, align: int = 32) -> Tuple[int, int]:
    """[begin, end) of the columns rank `rank` owns; shards are equal and `align`-column aligned (the decode
    kernel's slab width), so every rank launches the same kernels."""
    if n_features % world != 0 or (n_features // world) % align != 0:
        raise NotImplementedError(
            f"out_features={n_features} cannot be split into {world} column shards of a multiple of {align} columns")
    per = n_features // world
    return rank * per, (rank + 1) * per


def shard_tensors(W_q: Tensor, scales: Tensor, zeros: Tensor, bias: Optional[Tensor], n_features: int, rank: int,
                  world: int, contiguous: bool = True, mx: Optional[bool] = None):
    """Slice the tensors `get_tensor_args()` returns (+ bias) down to one column shard.

    W_q is `(K/eps, N)` (packed, or the `(K, N)` strided view of unpacked weights); scales / zeros are `(K/gs, N)`,
    `(1, N)` per-channel, MX scales `(N, K/gs)`, or the reference's placeholders (0-d scalar zero, empty `[[]]`),
    which are passed through untouched (core.py:470-475).

    `mx` says which axis of `scales` is the feature axis: MX layers store scales as the `(N, K/gs)` view with strides
    `(1, N)` (core.py:490-497), everything else as `(K/gs, N)`.  A shape test alone is ambiguous when K/gs == N
    (e.g. N = 256, K = 8192, gs = 32); with `mx=None` the storage order decides (strides `(1, N)` = MX view)."""
    b, e = column_range(n_features, rank, world)

    def cols(t: Tensor, feature_axis: int) -> Tensor:
        if t is None or t.dim() < 2 or t.numel() <= 1:
            return t                                     # scalar zero-point / empty placeholder
        if t.shape[feature_axis] != n_features:
            raise ValueError(f"cannot column-shard a tensor of shape {tuple(t.shape)} for N={n_features} "
                             f"along axis {feature_axis}")
        return t[b:e] if feature_axis == 0 else t[..., b:e]

    def meta_axis(t: Tensor) -> int:
        if t is None or t.dim() < 2 or t.numel() <= 1:
            return 1
        if mx is not None:
            return 0 if mx else 1
        if t.shape[0] == n_features and t.shape[-1] == n_features:
            return 0 if (t.stride(0) == 1 and t.stride(1) == n_features) else 1
        return 0 if (t.shape[0] == n_features and t.shape[-1] != n_features) else 1

    Wl = cols(W_q, 1)
    if contiguous and W_q.is_contiguous():
        Wl = Wl.contiguous()                             # packed (K/eps, N/g) row-major, what the kernels assume
    elif not W_q.is_contiguous():
        # unpacked 8/16-bit weights are the transposed view of an (N, K) row-major matrix: keep that storage order
        Wl = Wl.t().contiguous().t()
    sl, zl = cols(scales, meta_axis(scales)), cols(zeros, 1)
    if sl is not None and sl.dim() >= 2 and sl.numel() > 1:
        sl = sl.contiguous() if scales.is_contiguous() else sl.t().contiguous().t()
    if zl is not None and zl.dim() >= 2 and zl.numel() > 1:
        zl = zl.contiguous() if zeros.is_contiguous() else zl.t().contiguous().t()
    bl = None if bias is None else bias[b:e].contiguous()
    return Wl, sl, zl, bl


def gather_columns(y_local: Tensor, group=None) -> Tensor:
    """The exchange step: all-gather `(…, N/g)` shard outputs into `(…, N)` (rank-major column blocks)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return y_local
    y_local = y_local.contiguous()
    out = torch.empty((world,) + tuple(y_local.shape), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out.view(-1), y_local.view(-1), group=group)      # flat views: accepted by NCCL and gloo alike
    return out.movedim(0, -2).reshape(*y_local.shape[:-1], world * y_local.shape[-1])


class SymmetricOutputs:
    """Gathered outputs in NVLink-mapped symmetric memory: `rows` output rows of `n_global` columns, the same buffer on
    every rank (torch.distributed._symmetric_memory: cudaMalloc'd, exchanged over IPC, peer pointers valid in kernels).
    `ColumnShardedLinear.forward_scatter` makes the decode kernel store its columns into every rank's copy (the
    transfer is the kernel's own epilogue, 64 bytes per CTA and peer); `barrier()` is the only synchronisation of a step.
    B200-native replacement of the NCCL all-gather for M == 1; NCCL stays the fallback and the reported
    baseline."""

    def __init__(self, rows: int, n_global: int, dtype, device, group=None):
        import ctypes
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.rows, self.n_global, self.dtype = rows, n_global, dtype
        self.buf = symm_mem.empty(rows, n_global, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.world, self.rank = self.handle.world_size, self.handle.rank
        self._ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self._ctypes = ctypes
        self._row_arrays = {}

    def row_ptrs(self, row: int):
        """ctypes array of the address of `row` in every rank's buffer (own rank first: it is also the plain output)."""
        a = self._row_arrays.get(row)
        if a is None:
            off = row * self.n_global * self.buf.element_size()
            order = [self.rank] + [r for r in range(self.world) if r != self.rank]
            a = (self._ctypes.c_void_p * self.world)(*[self._ptrs[r] + off for r in order])
            self._row_arrays[row] = a
        return a

    def barrier(self):
        self.handle.barrier(channel=0)


class ColumnShardedLinear(torch.nn.Module):
    """One rank's shard of a packed `GemLiteLinear` + the all-gather of the outputs.

        layer  = A16W4_HQQ_INT(...).from_weights(...)             # full layer (any rank may build it)
        shard  = ColumnShardedLinear(layer, rank, world)          # keeps only columns [rank*N/world, ...)
        y      = shard(x)                                         # (…, N), identical on every rank
    """

    def __init__(self, layer, rank: int, world: int, group=None):
        super().__init__()
        from .core import GemLiteLinear
        from .dtypes import is_mx_dtype
        N = layer.out_features
        W_q, scales, zeros = layer.get_tensor_args()
        Wl, sl, zl, bl = shard_tensors(W_q.data, scales.data, zeros.data, None if layer.bias is None else layer.bias.data,
                                       N, rank, world, contiguous=bool(layer.data_contiguous),
                                       mx=is_mx_dtype(layer.input_dtype))
        n_local = N // world
        local = GemLiteLinear.__new__(GemLiteLinear)
        torch.nn.Module.__init__(local)
        meta = torch.tensor(layer.get_meta_args(), dtype=torch.int32)
        local.load_state_dict({"W_q": Wl, "bias": bl, "scales": sl, "zeros": zl, "metadata": meta,
                               "orig_shape": torch.tensor([n_local, layer.in_features], dtype=torch.int32)})
        local.forward = local.forward_auto_no_warmup
        self.local, self.rank, self.world, self.group = local, rank, world, group
        self.in_features, self.out_features = layer.in_features, N

    def forward_local(self, x: Tensor) -> Tensor:
        return self.local(x)

    def forward(self, x: Tensor) -> Tensor:
        return gather_columns(self.local(x), self.group)

    def forward_scatter(self, x: Tensor, out: "SymmetricOutputs", row: int = 0) -> None:
        """Decode call whose shard of the output is stored straight into row `row` of every rank's `out` buffer (fused
        output all-gather).  The caller brackets a step with `out.barrier()`.  NotImplementedError when the shard has no
        multi-destination kernel (anything but A16W4 at M == 1): use forward()."""
        from .core import matmul_bcast
        x2d = x.view(-1, x.shape[-1])
        matmul_bcast(x2d, self.local.bias, self.local.get_tensor_args(), self.local.get_meta_args(), out.row_ptrs(row),
                     self.rank * (self.out_features // self.world))
