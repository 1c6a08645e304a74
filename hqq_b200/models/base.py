"""Model walker and quantised-checkpoint I/O -- the immediate caller of ``HQQLinear``.

Mirrors the public surface of ``hqq/models/base.py`` (reference file:line cited per function) so that
``AutoHQQHFModel.quantize_model(model, quant_config, compute_dtype, device)``, ``save_quantized``, ``from_quantized`` and
``save_to_safetensors`` keep working when the import root is swapped (``hqq_b200.install_as_hqq()``):

* the walk replaces every ``nn.Linear`` (except what the per-tag config maps to ``None``) by an ``HQQLinear`` quantised on
  the GPU through ``libhqq_b200.so``;
* ``qmodel.pt`` is ``{module name: state_dict}`` with un-encoded HQQ state dicts, ``config.json`` the architecture --
  byte-compatible with checkpoints written by the reference;
* safetensors shards use the encoded (all-tensor) HQQ ``state_dict`` and the HF index format.

Host-side logic only: no arithmetic happens here.  Hub download is not provided (this framework has no network
dependency); ``from_quantized`` takes a local directory.
"""
from __future__ import annotations

import json
import os
from abc import abstractmethod
from contextlib import contextmanager
from typing import Callable, Iterable, Union

import torch
from torch import float16, nn

from ..core.quantize import HQQLinear
from ..core.utils import cleanup

# what counts as a "linear layer" for the walk (models/base.py:41-42); the reference also lists its LoRA / third-party
# backend wrappers, which are out of scope here
_QUANT_LAYERS = [nn.Linear, HQQLinear]
_IGNORE_LINEAR = ["lm_head"]


# ------------------------------------------------------------------------------------------- tree helpers (base.py:45-85)
def find_parent(model: nn.Module, name: str) -> nn.Module:
    """The module that owns the child called `name` (dotted path from `model`)."""
    parent = model
    for part in name.split(".")[:-1]:
        parent = parent._modules[part]
    return parent


def is_leaf_module(module: nn.Module) -> bool:
    return len(module._modules) == 0


def name_to_linear_tag(name: str) -> str:
    """``model.layers.31.self_attn.k_proj`` -> ``self_attn.k_proj``: the key of the per-layer-kind quant config."""
    return ".".join(p for p in name.split(".") if p not in ("model", "layers") and not p.isnumeric())


def get_all_children_from_model(model: nn.Module, ignore: Iterable[str] = ()) -> list:
    """Leaf module names in definition order."""
    ignore = set(ignore)
    return [name for name, module in model.named_modules() if is_leaf_module(module) and name.split(".")[-1] not in ignore]


def get_linear_tags_from_model(model: nn.Module, ignore: Iterable[str]) -> list:
    ignore = set(ignore)
    tags = []
    for name, module in model.named_modules():
        if type(module) in _QUANT_LAYERS and name.split(".")[-1] not in ignore:
            tag = name_to_linear_tag(name)
            if tag not in tags:
                tags.append(tag)
    return tags


def _tensors_to(value, device):
    return value.to(device) if isinstance(value, (torch.Tensor, nn.Parameter)) else value


def _hook_inputs_to_device(module: nn.Module, device) -> None:
    """Multi-device pipelines: tensors entering `module` are moved to its device first (base.py:88-106, 367-386)."""
    inner = module.forward
    module.device = device

    def forward(*args, **kwargs):
        return inner(*[_tensors_to(a, module.device) for a in args], **{k: _tensors_to(v, module.device) for k, v in kwargs.items()})

    module.forward_orig = inner
    module.forward = forward


def plan_device_map(all_nodes: list, all_blocks: Union[list, None], device) -> tuple:
    """Which device every leaf module lives on (base.py:297-336).  `device` is one device string, a list (blocks are split
    into contiguous runs, everything before the first block goes to the first device, everything after the last block to
    the last) or a ``{block name: device}`` dict.  Returns ``(device_map, num_devices)``."""
    if isinstance(device, dict):
        device_map = dict(device)
        all_blocks = list(device_map.keys())
        num_devices = len(set(device_map.values()))
    elif isinstance(device, (list, tuple)):
        devices = list(device)
        num_devices = len(devices)
        device_map = {}
        for node in all_nodes:
            if ".layers" in node:
                break
            device_map[node] = devices[0]
        for node in reversed(all_nodes):
            if ".layers" in node:
                break
            device_map[node] = devices[-1]
        blocks = all_blocks or []
        per_device = max(1, len(blocks) // num_devices)
        for j, block in enumerate(blocks):
            device_map[block] = devices[min(j // per_device, num_devices - 1)]
    else:
        device_map = {k: device for k in (all_blocks or []) + all_nodes}
        num_devices = 1
    blocks = all_blocks or []
    for node in all_nodes:
        owners = [b for b in blocks if b in node]
        device_map[node] = device_map[owners[-1] if owners else node]
    return device_map, num_devices


@contextmanager
def init_empty_weights():
    """Build a module tree whose parameters live on the meta device (buffers stay real), so an architecture can be created
    from its config without allocating the dense weights that ``from_quantized`` is about to replace."""
    register = nn.Module.register_parameter

    def register_on_meta(module, name, param):
        register(module, name, param)
        if param is not None:
            held = module._parameters[name]
            module._parameters[name] = nn.Parameter(held.to(torch.device("meta")), requires_grad=held.requires_grad)

    nn.Module.register_parameter = register_on_meta
    try:
        yield
    finally:
        nn.Module.register_parameter = register


# -------------------------------------------------------------------------------------------------- BasePatch (base.py:110-221)
class BasePatch:
    """How the layers of a model are visited and replaced.  Override `get_linear_tags` / `patch_*` for an architecture."""

    @classmethod
    def get_ignore_layers(cls, model) -> list:
        """Names that are containers, not layers: the root and every non-leaf module."""
        return [""] + [name for name, module in model.named_modules() if name and not is_leaf_module(module)]

    @classmethod
    def _replace_each(cls, model, names: list, make: Callable, verbose: bool) -> None:
        it = names
        if verbose:
            try:
                from tqdm import tqdm
                it = tqdm(names)
            except ImportError:
                pass
        for name in it:
            setattr(find_parent(model, name), name.split(".")[-1], make(name))
        cleanup()

    @classmethod
    def patch_nonlinearlayers(cls, model, patch_fct: Callable, verbose: bool = True) -> None:
        ignore = set(cls.get_ignore_layers(model))
        picked = {name: m for name, m in model.named_modules() if type(m) not in _QUANT_LAYERS and name not in ignore}
        cls._replace_each(model, list(picked), lambda name: patch_fct(picked[name]), verbose)

    @classmethod
    def patch_linearlayers(cls, model, patch_fct: Callable, patch_params: Union[dict, None], verbose: bool = True) -> None:
        ignore = set(cls.get_ignore_layers(model))
        picked = {name: m for name, m in model.named_modules() if type(m) in _QUANT_LAYERS and name not in ignore}
        params = patch_params or {}
        cls._replace_each(model, list(picked), lambda name: patch_fct(picked[name], params.get(name_to_linear_tag(name))), verbose)

    @classmethod
    def get_linear_tags(cls) -> list:
        return []

    @classmethod
    def set_auto_linear_tags(cls, model, ignore: list = _IGNORE_LINEAR) -> None:
        if not hasattr(model, "linear_tags"):
            tags = cls.get_linear_tags()
            model.linear_tags = tags if len(tags) > 0 else get_linear_tags_from_model(model, ignore=ignore)
            model.base_class = cls

    @classmethod
    def autoname_modules(cls, model) -> None:
        """Every module learns its dotted name: the key of its entry in a saved checkpoint."""
        for name, module in model.named_modules():
            module.name = name

    @classmethod
    def freeze_model(cls, model) -> None:
        for param in model.parameters():
            param.requires_grad = False

    @classmethod
    def patch_model(cls, model, patch_nonlinear_fct: Callable, patch_linear_fct: Callable, patch_params: dict, verbose: bool = True) -> None:
        model.eval()
        cls.freeze_model(model)
        cls.autoname_modules(model)
        cls.patch_nonlinearlayers(model, patch_nonlinear_fct, verbose=verbose)
        cls.patch_linearlayers(model, patch_linear_fct, patch_params, verbose=verbose)
        cleanup()


# ----------------------------------------------------------------------------------------------- BaseHQQModel (base.py:224-647)
class BaseHQQModel:
    """Quantise / save / load a whole model.  Combine with a `BasePatch` (see ``hqq_b200.models.hf.base``)."""

    @abstractmethod
    def create_model(cls, save_dir, kwargs):
        """An empty model of the saved architecture."""

    @abstractmethod
    def cache_model(cls, model, save_dir: str):
        """Write the architecture (no weights) to `save_dir`."""

    @classmethod
    def get_config_file(cls, save_dir: str) -> str:
        return os.path.join(save_dir, "config.json")

    @classmethod
    def get_weight_file(cls, save_dir: str) -> str:
        return os.path.join(save_dir, "qmodel.pt")

    @classmethod
    def save_weights(cls, weights: dict, save_dir: str) -> None:
        torch.save(weights, cls.get_weight_file(save_dir))

    @classmethod
    def load_weights(cls, save_dir: str, map_location=None):
        return torch.load(cls.get_weight_file(save_dir), map_location=map_location, weights_only=True)

    @classmethod
    def setup_model(cls, model) -> None:
        cls.autoname_modules(model)
        cls.set_auto_linear_tags(model)

    # -------------------------------------------------------------------------------------------- quantize (base.py:267-401)
    @classmethod
    def quantize_model(cls, model, quant_config: dict, compute_dtype: torch.dtype = float16, device: Union[str, list, dict] = "cuda"):
        """Replace the model's linear layers by `HQQLinear`s quantised on `device`.  `quant_config` is either one
        ``BaseQuantizeConfig`` for every layer or ``{linear tag: config or None}`` (tags that are not named stay dense)."""
        if getattr(model, "hqq_quantized", False):
            print("Model was already quantized")
            return None
        cls.setup_model(model)
        if any(key in model.linear_tags for key in quant_config.keys()):
            patch_params = {tag: None for tag in model.linear_tags}
            patch_params.update(quant_config)
        else:
            patch_params = {tag: quant_config for tag in model.linear_tags}

        all_nodes = get_all_children_from_model(model, [])
        try:
            layers = model.model.layers if hasattr(model, "model") else model.layers
            all_blocks = ["model.layers." + str(i) for i in range(len(layers))]
        except Exception:
            all_blocks = None
            if not isinstance(device, dict):
                print("Default model structure not supported. Make sure you feed device as dictionary as {name_block: device}")
        device_map, num_devices = plan_device_map(all_nodes, all_blocks, device)

        def patch_linear(layer, layer_config):
            if type(layer) is HQQLinear:
                return layer
            where = device_map[layer.name]
            if layer_config is not None:
                out = HQQLinear(layer, layer_config, compute_dtype=compute_dtype, device=where)
            else:
                out = layer.to(device=where, dtype=compute_dtype)
            out.device = where
            return out

        def patch_other(layer):
            where = device_map[layer.name]
            layer.device = where
            return layer.to(device=where, dtype=compute_dtype)

        cls.patch_model(model, patch_other, patch_linear, patch_params)

        if num_devices > 1:
            core = model if hasattr(model, "layers") else model.model
            _hook_inputs_to_device(getattr(core, all_nodes[0].split(".")[-1]), device_map[all_nodes[0]])
            for block in core.layers:
                _hook_inputs_to_device(block, device_map[block.name])
        model.base_class = cls
        model.hqq_quantized = True
        return model

    # ------------------------------------------------------------------------------------------------ save (base.py:405-432)
    @classmethod
    def serialize_weights(cls, model, verbose: bool = False) -> dict:
        """``{module name: state_dict}`` over the leaf modules; HQQ layers emit their un-encoded state dict (python
        scalars stay python scalars inside ``qmodel.pt``)."""
        weights = {}
        ignore = set(cls.get_ignore_layers(model))
        for name, module in model.named_modules():
            if name in ignore:
                continue
            try:
                module.encoded_state_dict = False
                state = module.state_dict()
                if len(state) > 0:
                    weights[name] = dict(state)
            except Exception:
                if verbose:
                    print("Skipping", name)
        return weights

    @classmethod
    def save_quantized(cls, model, save_dir: str, verbose: bool = False) -> None:
        os.makedirs(save_dir, exist_ok=True)
        cls.cache_model(model, save_dir)
        cls.save_weights(cls.serialize_weights(model, verbose=verbose), save_dir)

    # ------------------------------------------------------------------------------------------------ load (base.py:434-543)
    @classmethod
    def try_snapshot_download(cls, save_dir_or_hub: str, cache_dir: Union[str, None] = "") -> str:
        """Resolve a LOCAL checkpoint directory (under `cache_dir` when given).  The reference falls back to a hub download
        here; this framework does not reach the network, so a missing directory is an error."""
        save_dir = save_dir_or_hub if not cache_dir else os.path.join(cache_dir, save_dir_or_hub)
        if not os.path.exists(save_dir):
            raise FileNotFoundError(f"{save_dir}: no such checkpoint directory (hub download is not supported by hqq_b200)")
        if not os.path.exists(cls.get_weight_file(save_dir)):
            raise Exception("Weight file missing. Check your cache directory.")
        if not os.path.exists(cls.get_config_file(save_dir)):
            raise Exception("Config file missing. Check your cache directory.")
        return save_dir

    @classmethod
    def post_module_load(cls, model, weights: dict) -> None:
        """Hook for weights that belong to no module."""

    @classmethod
    def from_quantized(cls, save_dir_or_hub, compute_dtype: torch.dtype = float16, device="cuda", cache_dir: Union[str, None] = "",
                       adapter: Union[str, None] = None, **kwargs):
        if adapter is not None:
            raise NotImplementedError("hqq_b200: LoRA adapters (hqq/core/peft.py) are outside this framework's scope")
        save_dir = cls.try_snapshot_download(save_dir_or_hub, cache_dir)
        model = cls.create_model(save_dir, kwargs)
        model.save_dir = save_dir
        cls.setup_model(model)
        try:
            weights = cls.load_weights(save_dir, device)
        except Exception:
            print("Failed to load the weights")
            raise FileNotFoundError(cls.get_weight_file(save_dir))

        @torch.no_grad()
        def load_module(module, params=None):
            if module.name not in weights:
                return module.to(device=device, dtype=compute_dtype, non_blocking=True)
            state = weights[module.name]
            if "W_q" in state:
                name = module.name
                module = HQQLinear(linear_layer=None, quant_config=None, compute_dtype=compute_dtype, device=device)
                module.load_state_dict(state)
                module.name = name
            else:
                for key, value in state.items():
                    setattr(module, key, nn.Parameter(value.to(device=device, dtype=compute_dtype, non_blocking=True), requires_grad=False))
            return module

        cls.patch_model(model, load_module, load_module, {tag: None for tag in model.linear_tags})
        cls.post_module_load(model, weights)
        model.hqq_quantized = True
        model.base_class = cls
        return model

    # ----------------------------------------------------------------------------------------- safetensors (base.py:546-647)
    @classmethod
    def save_to_safetensors(cls, model, save_dir: str, num_blocks_per_file: int = 5, verbose: bool = True) -> None:
        """``config.json`` + the encoded state dict as ``model.safetensors`` or, for deep models, shards of
        `num_blocks_per_file` transformer blocks with an HF-style ``model.safetensors.index.json``."""
        from safetensors.torch import save_file

        def count_linears(module) -> int:
            n = 0
            for child in module.children():
                n += 1 if isinstance(child, (HQQLinear, nn.Linear)) else count_linears(child)
            return n

        config = getattr(model, "config", None)
        num_layers = config.num_hidden_layers if hasattr(config, "num_hidden_layers") else count_linears(model)
        os.makedirs(save_dir, exist_ok=True)
        if config is not None:
            if hasattr(config, "_attn_implementation_autoset"):
                del config._attn_implementation_autoset
            config.to_json_file(os.path.join(save_dir, "config.json"))

        for module in model.modules():  # safetensors holds tensors only (save_quantized switches the encoding off)
            if isinstance(module, HQQLinear):
                module.encoded_state_dict = True
        tensors = model.state_dict()
        host = lambda keys: {k: tensors[k].cpu().contiguous() for k in keys}
        num_chunks = num_layers // num_blocks_per_file
        if num_chunks <= 1:
            save_file(host(tensors.keys()), os.path.join(save_dir, "model.safetensors"))
            return
        total_size = sum(t.numel() * t.element_size() for t in tensors.values())
        files = [f"model-{i:05d}-of-{num_chunks:05d}.safetensors" for i in range(1, num_chunks + 1)]
        blocks_per_chunk = num_layers // num_chunks
        remaining = list(tensors.keys())
        weight_map = {}
        for chunk_id, fname in enumerate(files):
            if chunk_id == num_chunks - 1:
                keys = remaining
            else:
                marks = ["layers." + str(i) + "." for i in range(chunk_id * blocks_per_chunk, (chunk_id + 1) * blocks_per_chunk)]
                keys = [k for k in remaining if any(m in k for m in marks)]
            taken = set(keys)
            remaining = [k for k in remaining if k not in taken]
            if keys:
                if verbose:
                    print("saving", chunk_id + 1, ":", len(keys), "/", len(tensors))
                save_file(host(keys), os.path.join(save_dir, fname))
                weight_map.update({k: fname for k in keys})
        assert len(weight_map) == len(tensors)
        with open(os.path.join(save_dir, "model.safetensors.index.json"), "w") as fh:
            json.dump({"weight_map": weight_map, "metadata": {"total_size": total_size}}, fh)


__all__ = ["BasePatch", "BaseHQQModel", "find_parent", "is_leaf_module", "name_to_linear_tag", "get_all_children_from_model",
           "get_linear_tags_from_model", "plan_device_map", "init_empty_weights"]
