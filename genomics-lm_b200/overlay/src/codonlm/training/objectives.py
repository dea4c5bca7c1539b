"""`src.codonlm.training.objectives` served by the B200 implementation (reference: objectives.py:6-105)."""
import os
import sys

_PKG_PARENT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", ".."))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

from codonlm_b200.objectives import (DEFAULT_BOUNDARY_IDS, PAD_ID, multi_offset_lm_loss,  # noqa: E402,F401
                                     offset_target_mask, termination_aux_loss, termination_distance_bucket_labels)
