"""Golden vectors for `ProteinConditionalTransformer` from the UNMODIFIED reference (src/protein_lm/models.py:5-59):
seeded construction, embeddings scaled to a trained-like range, forward logits, next-token cross-entropy, backward.

    python tests/golden/make_protein_golden.py      (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.environ.get("CGPT_REFERENCE", "/root/reference"))
from src.protein_lm.config import ProteinLMConfig  # noqa: E402
from src.protein_lm.models import ProteinConditionalTransformer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(vocab_size=28, n_layer=2, n_head=4, n_embd=64, block_size=96, dropout=0.0)

torch.manual_seed(1337)
m = ProteinConditionalTransformer(ProteinLMConfig(**CFG))
import hashlib
h = hashlib.sha256()
for k, v in m.state_dict().items():
    h.update(k.encode())
    h.update(v.detach().numpy().tobytes())
g = torch.Generator().manual_seed(7)
with torch.no_grad():
    m.token_embedding.weight.mul_(0.3)
    m.position_embedding.weight.mul_(0.3)
    for name, p in m.named_parameters():
        if "norm" in name:
            p.add_(0.1 * torch.randn(p.shape, generator=g))
m.eval()
rng = np.random.default_rng(5)
idx = torch.from_numpy(rng.integers(1, 28, size=(3, 80)).astype(np.int64))
idx[1, 60:] = 0
logits = m(idx)
tgt = torch.roll(idx, -1, dims=1)
tgt[:, -1] = 0
loss = F.cross_entropy(logits.reshape(-1, 28), tgt.reshape(-1), ignore_index=0)
loss.backward()
out = {"idx": idx.numpy(), "targets": tgt.numpy(), "logits": logits.detach().numpy()}
for k, v in m.state_dict().items():
    out["sd." + k] = v.detach().numpy()
for k, p in m.named_parameters():
    out["grad." + k] = p.grad.detach().numpy()
out["meta"] = np.array(json.dumps(dict(cfg=CFG, loss=float(loss), init_sha256=h.hexdigest(), torch=torch.__version__)))
np.savez_compressed(os.path.join(HERE, "protein_lm.npz"), **out)
print("loss", float(loss), "logits absmax", float(logits.abs().max()), os.path.getsize(os.path.join(HERE, "protein_lm.npz")) // 1024, "KiB")
