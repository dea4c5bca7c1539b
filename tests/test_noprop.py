"""NoProp variant (reference model_tiny_gpt.py:391-459, tests/test_noprop.py): the oracle restatement
against golden vectors from the unmodified reference (tests/golden/make_noprop_golden.py), the module contract of the
CUDA implementation on CPU, and its numerics on the GPU."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import codon_gpt_oracle as O

GOLD = os.path.join(ROOT, "tests", "golden", "noprop.npz")


def _load():
    z = np.load(GOLD)
    ctor = ast.literal_eval(str(z["ctor"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.")}
    return z, ctor, sd, grads


def test_oracle_noprop_matches_reference_golden():
    z, ctor, sd, grads = _load()
    cfg = O.make_cfg(**{k: v for k, v in ctor.items()})
    leaves = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    leaves["head.weight"] = leaves["tok_emb.weight"]
    idx = torch.from_numpy(z["idx"])
    logits, preds = O.noprop_forward(leaves, cfg, idx, torch.from_numpy(z["target_embeddings"]))
    assert np.abs(logits.detach().numpy() - z["logits"]).max() <= 2e-5
    for l, p in enumerate(preds):
        assert np.abs(p.detach().numpy() - z[f"pred.{l}"]).max() <= 2e-5
    loss = logits.pow(2).mean() + sum(p.pow(2).mean() for p in preds)
    assert loss.item() == pytest.approx(float(z["loss"]), rel=3e-6)
    loss.backward()
    gmax = max(v.norm().item() for v in grads.values())
    for k, ref in grads.items():
        if k == "head.weight":
            continue
        got = leaves[k].grad
        assert (got - ref).norm().item() <= 2e-5 * ref.norm().item() + 1e-6 * gmax, k


def test_noprop_module_contract_on_cpu():
    from codonlm_b200.model_tiny_gpt import NoPropBlock, NoPropTinyGPT
    z, ctor, sd, _ = _load()
    torch.manual_seed(0)
    m = NoPropTinyGPT(**ctor)
    assert len(m.blocks) == ctor["n_layer"] and isinstance(m.blocks[0], NoPropBlock)  # reference tests/test_noprop.py:33-35
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("attn.mask")}
    assert mine == {k: tuple(v.shape) for k, v in sd.items()}
    assert m.head.weight is m.tok_emb.weight
    full = dict(sd)
    for l in range(ctor["n_layer"]):
        full[f"blocks.{l}.attn.mask"] = m.blocks[l].attn.mask
    m.load_state_dict(full, strict=True)
    from codonlm_b200 import _lib
    with pytest.raises(_lib.CgptError):  # no CPU path
        m(torch.from_numpy(z["idx"]))


@pytest.mark.gpu
def test_noprop_forward_backward_on_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from codonlm_b200.model_tiny_gpt import NoPropTinyGPT
    z, ctor, sd, grads = _load()
    m = NoPropTinyGPT(**ctor)
    full = dict(sd)
    for l in range(ctor["n_layer"]):
        full[f"blocks.{l}.attn.mask"] = m.blocks[l].attn.mask
    m.load_state_dict(full, strict=True)
    m = m.to("cuda").eval()
    idx = torch.from_numpy(z["idx"]).cuda()
    te = torch.from_numpy(z["target_embeddings"]).cuda()
    logits, preds = m(idx, target_embeddings=te)
    assert (logits.cpu() - torch.from_numpy(z["logits"])).abs().max().item() <= 2e-2
    for l, p in enumerate(preds):
        ref = torch.from_numpy(z[f"pred.{l}"])
        assert (p.float().cpu() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    loss = logits.float().pow(2).mean() + sum(p.float().pow(2).mean() for p in preds)
    assert loss.item() == pytest.approx(float(z["loss"]), rel=5e-3)
    loss.backward()
    e2 = n2 = 0.0
    for k, p in m.named_parameters():
        ref = grads[k].cuda()
        e2 += (p.grad.float() - ref).norm().item() ** 2
        n2 += ref.norm().item() ** 2
    assert e2 ** 0.5 <= 1e-2 * n2 ** 0.5
    logits2, preds2 = m(idx)  # without target embeddings (:409-410)
    assert len(preds2) == ctor["n_layer"] and logits2.shape == logits.shape
