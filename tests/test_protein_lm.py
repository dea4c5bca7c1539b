"""f4 — `ProteinConditionalTransformer` (src/protein_lm/models.py:5-59) on the CUDA kernels, against vectors from the
unmodified reference (tests/golden/make_protein_golden.py): constructor RNG contract and state-dict keys (CPU), logits /
loss / gradients (GPU)."""
import hashlib
import json
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import load_golden


def _load():
    z, _, sd, grads = load_golden("protein_lm")
    meta = json.loads(str(z["meta"]))
    return z, meta, sd, grads


def test_constructor_contract_and_state_dict_keys():
    from codonlm_b200.protein_lm import ProteinConditionalTransformer
    z, meta, sd, grads = _load()
    torch.manual_seed(1337)
    m = ProteinConditionalTransformer(SimpleNamespace(**meta["cfg"]))
    h = hashlib.sha256()
    for k, v in m.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().numpy().tobytes())
    assert h.hexdigest() == meta["init_sha256"]  # same modules in the same order: the reference's initial weights
    assert set(m.state_dict()) == set(sd)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(Exception):  # no CPU path
        m.eval()(torch.zeros((1, 4), dtype=torch.long))


@pytest.mark.gpu
def test_forward_backward_match_reference():
    from codonlm_b200.protein_lm import ProteinConditionalTransformer
    z, meta, sd, grads = _load()
    m = ProteinConditionalTransformer(SimpleNamespace(**meta["cfg"]))
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda").eval()
    idx = torch.from_numpy(z["idx"]).cuda()
    tgt = torch.from_numpy(z["targets"]).cuda()
    logits = m(idx)
    ref = torch.from_numpy(z["logits"]).cuda()
    scale = max(1.0, ref.abs().max().item() / 8.0)
    assert (logits - ref).abs().max().item() <= 2e-2 * scale
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), tgt.reshape(-1), ignore_index=0)
    assert loss.item() == pytest.approx(meta["loss"], rel=1e-3)
    loss.backward()
    gmax = max(v.norm().item() for v in grads.values())
    e2 = n2 = 0.0
    for name, p in m.named_parameters():
        r = grads[name].cuda()
        err, den = (p.grad - r).norm().item(), r.norm().item()
        e2, n2 = e2 + err * err, n2 + den * den
        assert err <= 1.25e-2 * max(den, 0.05 * gmax), (name, err, den)
    assert e2 ** 0.5 <= 1e-2 * n2 ** 0.5
    # a host-resident copy is staged for inference and returns host logits
    host = ProteinConditionalTransformer(SimpleNamespace(**meta["cfg"]))
    host.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = host.eval()(torch.from_numpy(z["idx"]))
    assert out.device.type == "cpu" and (out - logits.detach().cpu()).abs().max().item() <= 1e-5
