"""Hugging Face transformers flavour of the model walker (mirrors ``hqq/models/hf``)."""
