"""Overlay of the reference's `src.codonlm` package: same package, model module served by codonlm_b200.

`extend_path` appends the reference's own `src/codonlm` directory (found through the parent namespace package `src`),
so `src.codonlm.checkpoints`, `src.codonlm.generate`, ... are still the reference's files."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from .model_tiny_gpt import TinyGPT  # noqa: E402  (what the reference's own __init__ exports)

__all__ = ["TinyGPT"]
