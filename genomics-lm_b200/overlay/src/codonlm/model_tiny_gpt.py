"""`src.codonlm.model_tiny_gpt` served by the B200 implementation (reference: src/codonlm/model_tiny_gpt.py:1-461).

Every name the reference module defines is re-exported; callers (`checkpoints.build_codon_model_from_cfg`,
`scripts/query_model.py`, `scripts/extract_embeddings.py`, `training/loop.py`, the tests) import them unchanged."""
import os
import sys

_PKG_PARENT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

from codonlm_b200.model_tiny_gpt import (Block, CausalSelfAttention, NoPropBlock, NoPropTinyGPT,  # noqa: E402,F401
                                         RotaryEmbedding, SwiGLU, TinyGPT, apply_rotary_pos_emb, rotate_half)

__all__ = ["TinyGPT", "NoPropBlock", "NoPropTinyGPT"]
