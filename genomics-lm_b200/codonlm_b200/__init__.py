"""codonlm_b200 — B200-native (sm_100a) implementation of genomics-lm's codon-GPT step.

Host side of the drop-in: mirrors ``src.codonlm.model_tiny_gpt`` / ``src.codonlm.training.objectives``
of the reference and calls the C-ABI library ``libcgpt_b200.so`` (include/cgpt.h) for all arithmetic.
"""
from . import _lib  # noqa: F401
from .model_tiny_gpt import CausalSelfAttention, DecodeState, MaskSpec, TinyGPT  # noqa: F401
from .objectives import (multi_offset_lm_loss, offset_target_mask, termination_aux_loss,  # noqa: F401
                         termination_distance_bucket_labels, training_loss)

__all__ = ["TinyGPT", "CausalSelfAttention", "MaskSpec", "DecodeState", "multi_offset_lm_loss", "offset_target_mask",
           "termination_aux_loss", "termination_distance_bucket_labels", "training_loss"]
