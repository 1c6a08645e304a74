"""Small helpers mirroring hqq/core/utils.py:10-68 (cleanup, divisibility, the state_dict scalar codec)."""
from __future__ import annotations

import gc
import math
from typing import Union

import torch


def cleanup() -> None:
    try:
        torch.cuda.empty_cache()
    except Exception:
        pass
    gc.collect()


def is_divisible(val1: int, val2: int) -> bool:
    return int(val2 * math.ceil(val1 / val2)) == val1


def zero_pad_row(tensor: torch.Tensor, num_rows: int, dtype: Union[torch.dtype, None] = None) -> torch.Tensor:
    out = torch.zeros([num_rows, tensor.shape[1]], device=tensor.device, dtype=tensor.dtype if dtype is None else dtype)
    out[: len(tensor)] = tensor
    return out


# state_dict values must be tensors for safetensors: python scalars / str / dtype / Size are encoded as tensors
# with exactly the reference's conventions (utils.py:36-50) so checkpoints are interchangeable.
def encode_safetensor_type(data):
    if isinstance(data, (torch.Tensor, torch.nn.Parameter)):
        return data
    if isinstance(data, torch.Size):
        return torch.tensor(data)
    if isinstance(data, torch.dtype):
        data = str(data)
    if isinstance(data, bool):
        return torch.tensor(int(data), dtype=torch.uint8)
    if isinstance(data, int):
        return torch.tensor(data, dtype=torch.int32)
    if isinstance(data, float):
        return torch.tensor(data, dtype=torch.float32)
    if isinstance(data, str):
        return torch.tensor([ord(ch) for ch in data], dtype=torch.uint8)
    return None


_DTYPE_BY_NAME = {str(d): d for d in (torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.uint8, torch.int8,
                                       torch.int16, torch.int32, torch.int64, torch.bool)}


def decode_safetensor_type(data, data_type):
    if data_type in (torch.Tensor, torch.nn.Parameter):
        return data
    if data_type is torch.Size:
        return torch.Size(data)
    if data_type is bool:
        return bool(data.item())
    if data_type is int:
        return int(data.item())
    if data_type is float:
        return float(data.item())
    if data_type is str:
        return "".join(chr(int(i)) for i in data)
    if data_type is torch.dtype:
        name = "".join(chr(int(i)) for i in data)
        return _DTYPE_BY_NAME[name]  # no eval(): only known dtype names are accepted
    return data
