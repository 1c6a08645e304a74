"""``AutoHQQHFModel``: the generic walker for Hugging Face ``transformers`` models (mirrors ``hqq/models/hf/base.py:7-49``).

The architecture travels as the model's own ``config.json``; an empty model is rebuilt from it with its parameters on the
meta device (``hqq_b200.models.base.init_empty_weights`` -- the reference uses ``accelerate`` for this, which this image does
not ship) before ``from_quantized`` swaps the quantised layers in.
"""
from __future__ import annotations

from ..base import BaseHQQModel, BasePatch, init_empty_weights


def _auto_class_for(architectures):
    """The transformers auto class that rebuilds a saved architecture (hf/base.py:29-37: CausalLM and
    SequenceClassification heads are recognised, everything else is a bare AutoModel)."""
    import transformers

    names = list(architectures or [])
    if len(names) == 1:
        for marker, auto in (("CausalLM", transformers.AutoModelForCausalLM),
                             ("SequenceClassification", transformers.AutoModelForSequenceClassification)):
            if marker in names[0]:
                return auto
    return transformers.AutoModel


class BaseHQQHFModel(BaseHQQModel):
    """`cache_model` / `create_model` for transformers models: the architecture is the model's own config.json."""

    @classmethod
    def cache_model(cls, model, save_dir):
        config = model.config
        config.architectures = [type(model).__name__]
        config.save_pretrained(save_dir)

    @classmethod
    def create_model(cls, save_dir, kwargs):
        import transformers

        config = transformers.AutoConfig.from_pretrained(cls.get_config_file(save_dir))
        passthrough = {key: kwargs[key] for key in ("attn_implementation",) if key in kwargs}
        with init_empty_weights():
            return _auto_class_for(config.architectures).from_config(config, **passthrough)


class AutoHQQHFModel(BaseHQQHFModel, BasePatch):
    """Used when no architecture-specific patch class exists: linear tags are discovered from the model."""
