"""Bit-packing ops (device side) -- same call signatures as the reference's packers
(gemlite/bitpack.py:112-144 pack_weights_over_cols_triton, :146-205 unpack_over_cols), backed by the
CUDA kernels in csrc/gl_pack.cu through the C ABI (glb_pack_cols / glb_unpack_cols)."""
import torch

from . import _cabi
from .dtypes import PACKING_BITWIDTH_TO_TORCH_DTYPE


def _check(W_nbits: int, packing_bitwidth: int):
    assert packing_bitwidth in [8, 16, 32, 64], "Unsuported bitpacking width"
    assert W_nbits in [8, 4, 2, 1], "Unsuported nbits"


def pack_weights_over_cols(W_q: torch.Tensor, W_nbits: int, packing_bitwidth: int, transpose: bool):
    """(rows, cols) uint8 codes -> packed words along cols. Returns (tensor, elements_per_sample).

    With transpose=True the result is the (cols/eps, rows) view GemLiteLinear keeps (a transposed
    view of a contiguous (rows, cols/eps)... no: the kernel writes the transposed matrix directly, so the
    returned tensor is contiguous in its (cols/eps, rows) shape)."""
    _check(W_nbits, packing_bitwidth)
    _cabi.require_cuda(W_q, "pack_weights_over_cols")
    assert W_q.dtype == torch.uint8 and W_q.dim() == 2
    eps = packing_bitwidth // W_nbits
    rows, cols = W_q.shape
    assert cols % eps == 0, "number of columns must be divisible by elements_per_sample"
    W_q = W_q.contiguous()
    dt = PACKING_BITWIDTH_TO_TORCH_DTYPE[packing_bitwidth]
    shape = (cols // eps, rows) if transpose else (rows, cols // eps)
    out = torch.empty(shape, dtype=dt, device=W_q.device)
    with torch.cuda.device(W_q.device):
        rc = _cabi.load().glb_pack_cols(W_q.data_ptr(), rows, cols, W_nbits, packing_bitwidth, out.data_ptr(),
                                        int(bool(transpose)), _cabi.current_stream_ptr(W_q.device))
    _cabi.check(rc, "glb_pack_cols")
    return out, eps


def unpack_over_cols(W_q_packed: torch.Tensor, W_nbits: int, num_output_cols: int, dtype: torch.dtype = torch.uint8,
                     transpose: bool = False) -> torch.Tensor:
    """Inverse of pack_weights_over_cols. `W_q_packed` is (rows, cols/eps), or (cols/eps, rows) when
    transpose=True; returns (rows, num_output_cols) codes."""
    _cabi.require_cuda(W_q_packed, "unpack_over_cols")
    packing_bitwidth = W_q_packed.dtype.itemsize * 8
    _check(W_nbits, packing_bitwidth)
    W_q_packed = W_q_packed.contiguous()
    rows = W_q_packed.shape[1] if transpose else W_q_packed.shape[0]
    out = torch.empty((rows, num_output_cols), dtype=torch.uint8, device=W_q_packed.device)
    with torch.cuda.device(out.device):
        rc = _cabi.load().glb_unpack_cols(W_q_packed.data_ptr(), rows, num_output_cols, W_nbits, packing_bitwidth,
                                          out.data_ptr(), int(bool(transpose)), _cabi.current_stream_ptr(out.device))
    _cabi.check(rc, "glb_unpack_cols")
    return out if dtype == torch.uint8 else out.to(dtype)


# names the reference exports
pack_weights_over_cols_triton = pack_weights_over_cols
pack_weights_over_cols_cuda = pack_weights_over_cols
