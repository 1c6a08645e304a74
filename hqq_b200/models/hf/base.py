"""``AutoHQQHFModel``: the generic walker for Hugging Face ``transformers`` models (mirrors ``hqq/models/hf/base.py:7-49``).

The architecture travels as the model's own ``config.json``; an empty model is rebuilt from it with its parameters on the
meta device (``hqq_b200.models.base.init_empty_weights`` -- the reference uses ``accelerate`` for this, which this image does
not ship) before ``from_quantized`` swaps the quantised layers in.
"""
from __future__ import annotations

from ..base import BaseHQQModel, BasePatch, init_empty_weights


class BaseHQQHFModel(BaseHQQModel):
    @classmethod
    def cache_model(cls, model, save_dir):
        model.config.architectures = [model.__class__.__name__]
        model.config.save_pretrained(save_dir)

    @classmethod
    def create_model(cls, save_dir, kwargs):
        import transformers

        model_kwargs = {k: kwargs[k] for k in ("attn_implementation",) if k in kwargs}
        config = transformers.AutoConfig.from_pretrained(cls.get_config_file(save_dir))
        auto_class = transformers.AutoModel
        archs = config.architectures or []
        if len(archs) == 1:
            if "CausalLM" in archs[0]:
                auto_class = transformers.AutoModelForCausalLM
            elif "SequenceClassification" in archs[0]:
                auto_class = transformers.AutoModelForSequenceClassification
        with init_empty_weights():
            model = auto_class.from_config(config, **model_kwargs)
        return model


class AutoHQQHFModel(BaseHQQHFModel, BasePatch):
    """Used when no architecture-specific patch class exists: linear tags are discovered from the model."""
