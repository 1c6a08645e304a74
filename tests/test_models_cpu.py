"""CPU: model walker + checkpoint formats against fixtures written by the REAL reference
(tests/golden/make_golden_models.py): linear tags, device planning, qmodel.pt / safetensors layouts.  No arithmetic runs here:
loading a quantised checkpoint only moves tensors."""
import json
import os

import pytest
import torch

from hqq_b200.core.quantize import HQQLinear
from hqq_b200.models import base as mb
from hqq_b200.models.hf.base import AutoHQQHFModel

transformers = pytest.importorskip("transformers")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models")
INFO = json.load(open(os.path.join(GOLD, "info.json")))


def tiny_llama(layers=4):
    cfg = transformers.AutoConfig.from_pretrained(os.path.join(GOLD, "quantized", "config.json"))
    cfg.num_hidden_layers = layers
    torch.manual_seed(0)
    return transformers.LlamaForCausalLM(cfg)


def same(a, b, path=""):
    """Deep equality over dicts / tensors / python scalars (dtype and shape included)."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a.keys()) == sorted(b.keys()), (path, sorted(a.keys()), sorted(b.keys()))
        for k in a:
            same(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, torch.Tensor):
        assert isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a.cpu(), b.cpu()), path
    else:
        assert type(a) is type(b) and a == b, (path, a, b)


def test_linear_tags_follow_the_reference():
    model = tiny_llama()
    for name, tag in INFO["tags"].items():
        assert mb.name_to_linear_tag(name) == tag
    assert sorted(mb.get_linear_tags_from_model(model, ignore=["lm_head"])) == INFO["linear_tags"]
    AutoHQQHFModel.setup_model(model)
    assert sorted(model.linear_tags) == INFO["linear_tags"] and model.base_class is AutoHQQHFModel
    assert model.model.layers[1].mlp.up_proj.name == "model.layers.1.mlp.up_proj"
    assert mb.find_parent(model, "model.layers.2.self_attn.q_proj") is model.model.layers[2].self_attn
    leaves = mb.get_all_children_from_model(model)
    assert sorted(leaves) == sorted(INFO["device_list_map"].keys())
    ignore = AutoHQQHFModel.get_ignore_layers(model)
    assert "" in ignore and "model.layers.0.mlp" in ignore and "model.layers.0.mlp.up_proj" not in ignore


def test_device_plan_for_a_device_list_matches_the_reference():
    model = tiny_llama()
    nodes = mb.get_all_children_from_model(model)
    blocks = ["model.layers." + str(i) for i in range(4)]
    plan, n = mb.plan_device_map(nodes, blocks, ["cpu", "cpu:0"])
    assert n == 2
    assert {k: plan[k] for k in nodes} == INFO["device_list_map"]
    plan, n = mb.plan_device_map(nodes, blocks, "cuda:1")
    assert n == 1 and set(plan[k] for k in nodes) == {"cuda:1"}
    plan, n = mb.plan_device_map(nodes, None, {"model.layers.0": "a", "model.layers.1": "a", "model.layers.2": "b", "model.layers.3": "b",
                                               "model.embed_tokens": "a", "model.norm": "b", "model.rotary_emb": "b", "lm_head": "b"})
    assert n == 2 and plan["model.layers.1.mlp.up_proj"] == "a" and plan["model.layers.3.self_attn.q_proj"] == "b" and plan["lm_head"] == "b"
    # more devices than blocks: every block still gets a device
    plan, n = mb.plan_device_map(mb.get_all_children_from_model(tiny_llama(layers=2)), blocks[:2], ["d0", "d1", "d2"])
    assert plan["model.layers.0"] == "d0" and plan["model.layers.1"] == "d1"


def test_quantize_model_walks_like_the_reference(monkeypatch):
    """The walk itself (which layers are quantised with which config, on which device) with the GPU layer stubbed out."""
    made = {}

    class StubLinear(torch.nn.Module):
        def __init__(self, layer, cfg, compute_dtype=None, device=None):
            super().__init__()
            made[layer.name] = (cfg, compute_dtype, device)

    monkeypatch.setattr(mb, "HQQLinear", StubLinear)
    monkeypatch.setattr(mb, "_QUANT_LAYERS", [torch.nn.Linear, StubLinear])
    model = tiny_llama()
    cfg4, cfg2 = {"weight_quant_params": {"nbits": 4}}, {"weight_quant_params": {"nbits": 2}}
    out = AutoHQQHFModel.quantize_model(model, {"self_attn.q_proj": cfg4, "mlp.down_proj": cfg2}, compute_dtype=torch.float32, device="cpu")
    assert out is model and model.hqq_quantized and model.base_class is AutoHQQHFModel
    assert sorted(made) == sorted(f"model.layers.{i}.{t}" for i in range(4) for t in ("self_attn.q_proj", "mlp.down_proj"))
    assert made["model.layers.3.mlp.down_proj"] == (cfg2, torch.float32, "cpu")
    assert type(model.model.layers[0].self_attn.k_proj) is torch.nn.Linear and type(model.lm_head) is torch.nn.Linear  # unnamed tags stay dense
    assert type(model.model.layers[0].self_attn.q_proj) is StubLinear
    assert not any(p.requires_grad for p in model.parameters())
    assert AutoHQQHFModel.quantize_model(model, cfg4) is None  # second call is a no-op, as in the reference
    # one config for every tag (lm_head is never a tag)
    made.clear()
    model = tiny_llama(layers=2)
    AutoHQQHFModel.quantize_model(model, cfg4, compute_dtype=torch.float32, device=["cpu", "cpu:0"])
    assert len(made) == 2 * 7 and "lm_head" not in made
    assert made["model.layers.0.mlp.up_proj"][2] == "cpu" and made["model.layers.1.mlp.up_proj"][2] == "cpu:0"
    assert hasattr(model.model.layers[1], "forward_orig") and hasattr(model.model.embed_tokens, "forward_orig")  # device hand-over hooks


def load_reference_checkpoint():
    return AutoHQQHFModel.from_quantized(os.path.join(GOLD, "quantized"), compute_dtype=torch.float32, device="cpu", cache_dir=None)


def test_from_quantized_reads_a_reference_checkpoint_and_writes_it_back(tmp_path):
    model = load_reference_checkpoint()
    assert model.hqq_quantized and model.base_class is AutoHQQHFModel
    types = {n: type(m).__name__ for n, m in model.named_modules() if len(m._modules) == 0}
    assert types == INFO["module_types"]
    ref = torch.load(os.path.join(GOLD, "quantized", "qmodel.pt"), weights_only=True)
    layer = model.model.layers[2].mlp.up_proj
    assert isinstance(layer, HQQLinear) and layer.meta["nbits"] == 3 and layer.W_q.dtype == torch.int32 and layer.ready
    assert (layer.in_features, layer.out_features) == (64, 128)
    assert torch.equal(layer.W_q.data, ref["model.layers.2.mlp.up_proj"]["W_q"])
    assert model.model.layers[1].mlp.down_proj.meta["axis"] == 0
    assert not any(p.is_meta for p in model.parameters())
    # our serialisation of the loaded model is the reference's file, entry by entry
    same(AutoHQQHFModel.serialize_weights(model), {k: dict(v) for k, v in ref.items()})
    out = str(tmp_path / "resaved")
    AutoHQQHFModel.save_quantized(model, out)
    same(torch.load(os.path.join(out, "qmodel.pt"), weights_only=True), ref)
    assert json.load(open(os.path.join(out, "config.json")))["architectures"] == ["LlamaForCausalLM"]
    again = AutoHQQHFModel.from_quantized(out, compute_dtype=torch.float32, device="cpu", cache_dir=None)
    same(AutoHQQHFModel.serialize_weights(again), {k: dict(v) for k, v in ref.items()})


def test_missing_checkpoint_parts_raise(tmp_path):
    with pytest.raises(FileNotFoundError):
        AutoHQQHFModel.from_quantized(str(tmp_path / "nope"), device="cpu", cache_dir=None)
    os.makedirs(tmp_path / "half")
    with pytest.raises(Exception, match="Weight file missing"):
        AutoHQQHFModel.from_quantized(str(tmp_path / "half"), device="cpu", cache_dir=None)


@pytest.mark.parametrize("blocks_per_file,gold", [(5, "st_single"), (2, "st_sharded")])
def test_safetensors_export_matches_the_reference_files(tmp_path, blocks_per_file, gold):
    from safetensors.torch import load_file
    model = load_reference_checkpoint()
    AutoHQQHFModel.serialize_weights(model)  # leaves the layers in un-encoded mode, like save_quantized; the export must not care
    out = str(tmp_path / gold)
    AutoHQQHFModel.save_to_safetensors(model, out, num_blocks_per_file=blocks_per_file, verbose=False)
    gdir = os.path.join(GOLD, gold)
    assert sorted(os.listdir(out)) == sorted(os.listdir(gdir))
    for f in os.listdir(gdir):
        if f.endswith(".safetensors"):
            same(load_file(os.path.join(out, f)), load_file(os.path.join(gdir, f)), f)
        elif f.endswith(".index.json"):
            assert json.load(open(os.path.join(out, f))) == json.load(open(os.path.join(gdir, f)))
    # and the encoded state dict loads back into fresh layers through nn.Module.load_state_dict
    tensors = {}
    for f in os.listdir(out):
        if f.endswith(".safetensors"):
            tensors.update(load_file(os.path.join(out, f)))
    prefix = "model.layers.0.self_attn.q_proj."
    fresh = HQQLinear(None, None, compute_dtype=torch.float32, device="cpu")
    fresh.load_state_dict({k[len(prefix):]: v for k, v in tensors.items() if k.startswith(prefix)})
    assert torch.equal(fresh.W_q.data, model.model.layers[0].self_attn.q_proj.W_q.data) and fresh.meta["nbits"] == 4


def test_install_as_hqq_exposes_the_model_modules():
    import hqq_b200
    hqq_b200.install_as_hqq()
    from hqq.models.base import BaseHQQModel, BasePatch  # noqa: F401
    from hqq.models.hf.base import AutoHQQHFModel as A
    assert A is AutoHQQHFModel
