"""Stand-in for the reference's native ``hqq_aten`` extension module (hqq/kernels/hqq_aten_cuda.cpp:32-73), served by
``libhqq_b200.so``.  Same function names and argument order; unlike the reference's kernels, ``dequantize`` also
handles ``axis=1`` (the reference asserts axis == 0, hqq_aten_cuda.cpp:35).

    import hqq_b200.hqq_aten as hqq_aten
    W = hqq_aten.dequantize(W_q, scale, zero, N, K, group_size, nbits, axis, packing)
"""
from __future__ import annotations

import torch

from . import ops

_BITS = {"8bit_u8": 8, "4bit_u8": 4, "3bit_32": 3, "2bit_u8": 2, "1bit_u8": 1}


def dequantize(W_q: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor, N: int, K: int, group_size: int, nbits: int, axis: int,
               packing: str) -> torch.Tensor:
    """-> Tensor[N, K] in scale.dtype (the fake/meta registration of the reference's custom op, quantize.py:261-263)."""
    if group_size is None or group_size <= 0:  # the reference passes -1 for "no grouping": one group per row / column
        group_size = K if axis == 1 else N
    return ops.dequantize(W_q, scale, zero, (N, K), group_size, _BITS[packing], axis, scale.dtype)


def _dq(nbits):
    def f(Wq_packed, scale, zero):
        # per-bit entry points work on the grouped matrix with axis-0 meta [1, w], like the reference kernels
        prow, w = Wq_packed.shape
        rows = prow * ops.FIELDS[nbits]
        out = ops.dequantize(Wq_packed, scale, zero, (rows, w), rows, nbits, 0, scale.dtype) if nbits != 3 else \
            _dq3(Wq_packed, scale, zero)
        return out
    return f


def _dq3(Wq_packed, scale, zero):
    W_r = ops.unpack(Wq_packed, 3, scale.dtype)  # keeps the padded rows, as the reference kernel does
    return (W_r - zero) * scale


def unpack_4bit_u8(Wq_packed):
    return ops.unpack(Wq_packed, 4, torch.uint8)


def unpack_2bit_u8(Wq_packed):
    return ops.unpack(Wq_packed, 2, torch.uint8)


def unpack_1bit_u8(Wq_packed):
    return ops.unpack(Wq_packed, 1, torch.uint8)


def unpack_3bit_32(Wq_packed):
    return ops.unpack(Wq_packed, 3, torch.uint8)


dequantize_8bit_u8 = _dq(8)
dequantize_4bit_u8 = _dq(4)
dequantize_2bit_u8 = _dq(2)
dequantize_1bit_u8 = _dq(1)
dequantize_3bit_32 = _dq3
