"""codonlm_b200 — B200-native (sm_100a) implementation of genomics-lm's codon-GPT step.

Host side of the drop-in: mirrors ``src.codonlm.model_tiny_gpt`` / ``src.codonlm.training.objectives``
of the reference and calls the C-ABI library ``libcgpt_b200.so`` (include/cgpt.h) for all arithmetic.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
