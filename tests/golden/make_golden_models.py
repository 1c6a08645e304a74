"""Golden fixtures for the model walker / checkpoint formats, produced by the REAL reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_models.py

A tiny random-init ``LlamaForCausalLM`` (transformers) is quantised on the CPU by the reference's
``BaseHQQModel + BasePatch`` (hqq/models/base.py) and written out with the reference's own ``save_quantized`` and
``save_to_safetensors``; the reference's device planning for a two-device list is recorded as well.  Stored under
``tests/golden/models/``.  ``hqq.models.hf`` needs ``accelerate`` (absent here), so the two abstract methods are filled in
exactly as hqq/models/hf/base.py:10-15 does (config.save_pretrained).
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "models")
REF = os.environ.get("HQQ_REFERENCE", "/root/reference")

shim = tempfile.mkdtemp()
with open(os.path.join(shim, "termcolor.py"), "w") as f:
    f.write("def colored(s, *a, **k):\n    return s\n")
sys.path.insert(0, shim)
sys.path.insert(0, REF)

import torch  # noqa: E402
import transformers  # noqa: E402
from hqq.core.quantize import BaseQuantizeConfig, HQQLinear  # noqa: E402
from hqq.models.base import BaseHQQModel, BasePatch, name_to_linear_tag  # noqa: E402


class RefAuto(BaseHQQModel, BasePatch):
    @classmethod
    def cache_model(cls, model, save_dir):
        model.config.architectures = [model.__class__.__name__]
        model.config.save_pretrained(save_dir)


def tiny_llama(seed=0, layers=4):
    torch.manual_seed(seed)
    cfg = transformers.LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=layers, num_attention_heads=4,
                                   num_key_value_heads=2, vocab_size=96, max_position_embeddings=64, tie_word_embeddings=False)
    return transformers.LlamaForCausalLM(cfg)


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    # the dense weights the model was quantised from (our side re-quantises them on the GPU and compares)
    model = tiny_llama()
    torch.save({k: v.clone() for k, v in model.state_dict().items()}, os.path.join(OUT, "dense_state_dict.pt"))
    quant_config = {"self_attn.q_proj": BaseQuantizeConfig(nbits=4, group_size=64, axis=1),
                    "self_attn.k_proj": BaseQuantizeConfig(nbits=4, group_size=64, axis=1),
                    "self_attn.v_proj": BaseQuantizeConfig(nbits=8, group_size=64, axis=1),
                    "self_attn.o_proj": None,  # stays dense
                    "mlp.gate_proj": BaseQuantizeConfig(nbits=2, group_size=64, axis=1),
                    "mlp.up_proj": BaseQuantizeConfig(nbits=3, group_size=64, axis=1),
                    "mlp.down_proj": BaseQuantizeConfig(nbits=4, group_size=64, axis=0)}
    RefAuto.quantize_model(model, quant_config=quant_config, compute_dtype=torch.float32, device="cpu")
    info = {"linear_tags": sorted(model.linear_tags),
            "module_types": {n: type(m).__name__ for n, m in model.named_modules() if len(m._modules) == 0},
            "tags": {n: name_to_linear_tag(n) for n, m in model.named_modules() if isinstance(m, (HQQLinear, torch.nn.Linear))}}
    # 1. qmodel.pt + config.json written by the reference
    qdir = os.path.join(OUT, "quantized")
    os.makedirs(qdir)
    RefAuto.save_quantized(model, qdir)
    # 2. safetensors: single file and 2 shards (4 blocks, 2 per file)
    for m in model.modules():
        m.encoded_state_dict = True  # save_quantized switched the encoding off
    RefAuto.save_to_safetensors(model, os.path.join(OUT, "st_single"), num_blocks_per_file=5, verbose=False)
    RefAuto.save_to_safetensors(model, os.path.join(OUT, "st_sharded"), num_blocks_per_file=2, verbose=False)
    # 3. device planning for a list of two (distinguishable) CPU devices
    model2 = tiny_llama()
    RefAuto.quantize_model(model2, quant_config=BaseQuantizeConfig(nbits=4, group_size=64), compute_dtype=torch.float32, device=["cpu", "cpu:0"])
    info["device_list_map"] = {n: str(m.device) for n, m in model2.named_modules() if len(m._modules) == 0}
    info["device_list_linear_tags"] = sorted(model2.linear_tags)
    with open(os.path.join(OUT, "info.json"), "w") as fh:
        json.dump(info, fh, indent=1, sort_keys=True)
    for root, _, files in os.walk(OUT):
        for f in files:
            p = os.path.join(root, f)
            print(os.path.relpath(p, OUT), os.path.getsize(p))


if __name__ == "__main__":
    main()
