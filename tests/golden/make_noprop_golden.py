"""Golden vectors for the NoProp variant, produced by the UNMODIFIED reference (run in the build container):
  python tests/golden/make_noprop_golden.py
Writes tests/golden/noprop.npz: weights of a seeded NoPropTinyGPT (model_tiny_gpt.py:418-459), token ids with <SEP>
boundaries, target embeddings, the logits and per-block predictions of forward(idx, target_embeddings), and the
gradients of  mean(logits^2) + sum_l mean(pred_l^2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.environ.get("CGPT_REFERENCE", "/root/reference"))
from src.codonlm.model_tiny_gpt import NoPropTinyGPT  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.codon_gpt_oracle import synthetic_batch  # noqa: E402

ctor = dict(vocab_size=68, block_size=48, n_layer=2, n_head=2, n_embd=64, dropout=0.0, sep_id=3, use_sdpa=True)
torch.manual_seed(1337)
m = NoPropTinyGPT(**ctor).eval()
g = torch.Generator().manual_seed(3)
with torch.no_grad():
    m.tok_emb.weight.mul_(0.05)
    m.pos_emb.weight.mul_(0.05)
    for name, p in m.named_parameters():
        if ".ln" in name or name.startswith("ln_f"):
            p.add_(0.1 * torch.randn(p.shape, generator=g))
idx, _ = synthetic_batch(3, 40, seed=21, realistic=True)
idx[:, 17] = 3
tgt_emb = 0.1 * torch.randn((3, 40, 64), generator=g)
logits, preds = m(idx, target_embeddings=tgt_emb)
loss = logits.pow(2).mean() + sum(p.pow(2).mean() for p in preds)
loss.backward()
out = {"idx": idx.numpy(), "target_embeddings": tgt_emb.numpy(), "logits": logits.detach().numpy(),
       "loss": np.array(float(loss)), "ctor": np.array(repr(ctor))}
for l, p in enumerate(preds):
    out[f"pred.{l}"] = p.detach().numpy()
for k, v in m.state_dict().items():
    if not k.endswith("attn.mask"):
        out["sd." + k] = v.detach().numpy()
for k, p in m.named_parameters():
    out["grad." + k] = p.grad.detach().numpy()
np.savez_compressed(os.path.join(HERE, "noprop.npz"), **out)
print("wrote noprop.npz, loss", float(loss), "keys", len(out))
