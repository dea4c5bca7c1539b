"""Shared pytest setup: import paths, the ``gpu`` marker, golden-fixture loader."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "genomics-lm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["gelu_abs_sep", "gelu_default_init", "swiglu_rope_causal", "gqa_untied_weighted",
                "heads_offsets_term", "window5", "shape_guidance"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(case):
    import torch
    z = np.load(os.path.join(GOLDEN_DIR, f"{case}.npz"))
    meta = json.loads(str(z["meta"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.")}
    return z, meta, sd, grads


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return (request.param,) + load_golden(request.param)
