"""Host-side helpers behind the names of ``hqq/core/utils.py`` (``cleanup``, ``is_divisible``, ``zero_pad_row`` and the
``state_dict`` scalar codec of utils.py:36-68).

The codec is a wire format shared with the reference -- safetensors stores tensors only, so the python values inside an
``HQQLinear.state_dict()`` travel as small tensors: bool -> uint8 scalar, int -> int32 scalar, float -> float32 scalar,
str -> uint8 code points, torch.dtype -> its ``str()`` as text, torch.Size -> int64 vector.  It is written here as two lookup
tables; unknown dtype names are rejected instead of evaluated.
"""
from __future__ import annotations

import gc
from typing import Optional

import torch


def cleanup() -> None:
    """Drop python garbage and hand cached device blocks back to the driver."""
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def is_divisible(val1: int, val2: int) -> bool:
    """`val2` divides `val1` (group sizes against tensor sizes)."""
    return val2 != 0 and val1 % val2 == 0


def zero_pad_row(tensor: torch.Tensor, num_rows: int, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """`tensor` on top of zero rows, `num_rows` rows in total (3-bit packing pads the row count to a multiple of 10)."""
    padded = tensor.new_zeros((num_rows, tensor.shape[1]), dtype=dtype or tensor.dtype)
    padded[: tensor.shape[0]].copy_(tensor)
    return padded


# ------------------------------------------------------------------------------------------------------ state_dict scalar codec
def _text_to_tensor(text: str) -> torch.Tensor:
    return torch.tensor([ord(ch) for ch in text], dtype=torch.uint8)


def _tensor_to_text(codes: torch.Tensor) -> str:
    return "".join(map(chr, codes.tolist()))


# order matters: bool is an int, Parameter is a Tensor
_ENCODERS = (
    (torch.Tensor, lambda v: v),
    (torch.Size, lambda v: torch.tensor(tuple(v))),
    (torch.dtype, lambda v: _text_to_tensor(str(v))),
    (bool, lambda v: torch.tensor(int(v), dtype=torch.uint8)),
    (int, lambda v: torch.tensor(v, dtype=torch.int32)),
    (float, lambda v: torch.tensor(v, dtype=torch.float32)),
    (str, _text_to_tensor),
)

_DTYPES = {str(d): d for d in (torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.uint8, torch.int8, torch.int16,
                               torch.int32, torch.int64, torch.bool)}

_DECODERS = {
    torch.Tensor: lambda t: t,
    torch.nn.Parameter: lambda t: t,
    torch.Size: lambda t: torch.Size(t.tolist()),
    bool: lambda t: bool(t.item()),
    int: lambda t: int(t.item()),
    float: lambda t: float(t.item()),
    str: _tensor_to_text,
    torch.dtype: lambda t: _DTYPES[_tensor_to_text(t)],
}


def encode_safetensor_type(data):
    """A python value of an HQQ state dict as the tensor that represents it on disk (None for unsupported types)."""
    for kind, encode in _ENCODERS:
        if isinstance(data, kind):
            return encode(data)
    return None


def decode_safetensor_type(data, data_type):
    """Inverse of `encode_safetensor_type` for a value whose python type is `data_type` (quantize.py `_META_TYPE`)."""
    decode = _DECODERS.get(data_type)
    return data if decode is None else decode(data)
