"""CPU: host-side mirror of hqq.core.quantize (config dicts, enums, codec, shapes, loud failure without CUDA)."""
import numpy as np
import pytest
import torch

from hqq_b200.core.quantize import BaseQuantizeConfig, HQQBackend, HQQLinear, Quantizer, hqq_base_quant_config
from hqq_b200.core.utils import decode_safetensor_type, encode_safetensor_type, is_divisible
from hqq_b200 import ops


def test_base_quantize_config_matches_reference_dict():
    cfg = BaseQuantizeConfig(nbits=4, group_size=64)
    assert cfg == {"weight_quant_params": {"nbits": 4, "channel_wise": True, "group_size": 64, "optimize": True,
                                           "round_zero": True, "axis": 1, "view_as_float": False},
                   "scale_quant_params": None, "zero_quant_params": None, "offload_meta": False}
    assert BaseQuantizeConfig(nbits=2)["weight_quant_params"]["round_zero"] is False  # round_zero only for 4-bit
    assert BaseQuantizeConfig is hqq_base_quant_config
    with pytest.raises(AssertionError):
        BaseQuantizeConfig(nbits=7)
    with pytest.raises(AssertionError):
        BaseQuantizeConfig(nbits=4, group_size=12)
    c = BaseQuantizeConfig(nbits=4, quant_zero=True, quant_scale=True)
    assert c["scale_quant_params"] == {"nbits": 8, "channel_wise": True, "group_size": 128, "optimize": False}
    assert c["zero_quant_params"] == {"nbits": 8, "channel_wise": False, "group_size": None, "optimize": False}


def test_backend_enum_and_registries():
    assert HQQBackend.PYTORCH.value == "forward_pytorch_backprop"
    assert HQQBackend.ATEN.value == "forward_aten_backprop"
    assert {m.name for m in HQQBackend} >= {"PYTORCH", "PYTORCH_COMPILE", "ATEN", "PYTORCH_FORWARD", "ATEN_FORWARD_INT8"}
    for m in HQQBackend:
        assert callable(getattr(HQQLinear, m.value))
    HQQLinear.set_backend(HQQBackend.ATEN)
    assert HQQLinear.backend is HQQBackend.ATEN
    HQQLinear.set_backend(HQQBackend.PYTORCH)
    assert Quantizer.SUPPORTED_BITS == [8, 6, 5, 4, 3, 2, 1.58, 1]
    assert Quantizer.bit_to_packing[3] == "3bit_32" and Quantizer.unpack_view_dtype["3bit_32"] == torch.int32
    assert set(Quantizer.pack) == set(Quantizer.unpack) == {"8bit_u8", "4bit_u8", "3bit_32", "2bit_u8", "1bit_u8"}


def test_state_dict_codec_matches_reference_encoding(golden):
    sd = golden.state_dict
    assert np.array_equal(encode_safetensor_type(4).numpy(), sd["sd/nbits"]) and encode_safetensor_type(4).dtype == torch.int32
    assert np.array_equal(encode_safetensor_type(True).numpy(), sd["sd/optimize"])
    assert np.array_equal(encode_safetensor_type("4bit_u8").numpy(), sd["sd/packing"])
    assert np.array_equal(encode_safetensor_type(torch.float32).numpy(), sd["sd/compute_dtype"])
    assert np.array_equal(encode_safetensor_type(torch.Size([64, 64])).numpy(), sd["sd/shape"])
    for v, t in ((7, int), (True, bool), (0.5, float), ("4bit_u8", str), (torch.bfloat16, torch.dtype), (torch.Size([3, 5]), torch.Size)):
        assert decode_safetensor_type(encode_safetensor_type(v), t) == v
    layer = HQQLinear(None, None, initialize=False)
    assert layer.state_dict_keys() == {k[3:] for k in sd.files} | {"bias"}
    assert all(v is None for v in layer.state_dict().values())


def test_packed_shapes():
    assert ops.packed_shape(4096, 4096, 64, 4, 1) == ((131072, 64), (262144, 64), 262144)
    assert ops.packed_shape(128, 256, 64, 3, 1)[0] == (52, 64)  # 512 rows -> ceil(512/10) int32 rows
    assert ops.packed_shape(128, 256, 64, 4, 0)[0] == (32, 512)
    assert is_divisible(128, 64) and not is_divisible(100, 64)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from hqq_b200.core.bitpack import BitPack
    with pytest.raises(RuntimeError, match="no CPU path"):
        BitPack.pack_4bit_u8(torch.zeros(8, 8, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        HQQLinear(torch.nn.Linear(64, 64), BaseQuantizeConfig(nbits=4, group_size=64), device="cuda")
