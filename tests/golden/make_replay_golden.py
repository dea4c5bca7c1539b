"""Golden vectors for the REPLAY term of the trainer's loss composition (src/codonlm/training/loop.py:1113-1141), from
the UNMODIFIED reference: its TinyGPT, its objectives, and the very statements of fwd() — main forward with aux, offset
and termination losses, then the second forward over the replay batch, `termination_aux_loss` on its sparse labels with
`replay_class_weights`, `total += replay_loss_weight * replay`.  The replay batch is built by the reference's own
`GeneratedTerminationReplayDataset` (src/codonlm/replay.py) from a JSONL written here.

    python tests/golden/make_replay_golden.py      (build container only: needs /root/reference)
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

REF = os.environ.get("CGPT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from src.codonlm.model_tiny_gpt import TinyGPT  # noqa: E402
from src.codonlm.replay import GeneratedTerminationReplayDataset  # noqa: E402
from src.codonlm.training import objectives as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.codon_gpt_oracle import synthetic_batch  # noqa: E402  (token generator only)

CTOR = dict(vocab_size=68, block_size=64, n_layer=2, n_head=2, n_embd=64, dropout=0.0, label_smoothing=0.05, sep_id=3,
            use_sdpa=True, termination_aux=True, multi_offset_targets=[2, 4])
OW = {2: 0.5, 4: 0.25}
TW, RW = 0.3, 0.2
REPLAY_CW = [1.0, 2.0, 0.5, 1.5, 1.0]

torch.manual_seed(1337)
m = TinyGPT(**CTOR)
g = torch.Generator().manual_seed(7)
with torch.no_grad():
    m.tok_emb.weight.mul_(0.02)
    m.pos_emb.weight.mul_(0.02)
    for name, p in m.named_parameters():
        if ".ln" in name or name.startswith("ln_f"):
            p.add_(0.1 * torch.randn(p.shape, generator=g))
        if name.startswith("offset_projs") or name.startswith("termination_head"):
            p.add_(0.02 * torch.randn(p.shape, generator=g))
m.eval()
idx, tgt = synthetic_batch(3, 64, seed=11, realistic=True)

# replay records: generated contexts with sparse termination labels, through the reference's dataset class
rng = np.random.default_rng(3)
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "replay.jsonl")
    with open(path, "w") as f:
        for r in range(4):
            n = int(rng.integers(20, 90))  # some longer than block_size: left-clipped by the dataset
            ids = [1] + [int(v) for v in rng.integers(4, 68, size=n - 1)]
            labels = [{"pos": int(p_), "class": int(rng.integers(0, 5))} for p_ in sorted(rng.choice(n, size=5, replace=False))]
            f.write(json.dumps({"ids": ids, "labels": labels}) + "\n")
    ds = GeneratedTerminationReplayDataset(path, block_size=64)
    replay_x = torch.stack([ds[i][0] for i in range(len(ds))])
    replay_labels = torch.stack([ds[i][1] for i in range(len(ds))])

# ---- the statements of fwd(), loop.py:1078-1142
logits_, next_loss_, aux_ = m(idx, tgt, return_aux=True)
total = next_loss_
off_total, off_losses = O.multi_offset_lm_loss(aux_["offset_logits"], tgt, OW, label_smoothing=CTOR["label_smoothing"],
                                               loss_weights=None)
total = total + off_total
labels = O.termination_distance_bucket_labels(tgt, stop_ids=(2,), bucket_edges=(0, 3, 10, 30))
term = O.termination_aux_loss(aux_["termination_logits"], labels, class_weights=None)
total = total + TW * term
_, _, replay_aux = m(replay_x, return_aux=True)
replay_loss = O.termination_aux_loss(replay_aux["termination_logits"], replay_labels,
                                     class_weights=torch.tensor(REPLAY_CW))
total = total + RW * replay_loss
total.backward()

out = {"idx": idx.numpy(), "targets": tgt.numpy(), "replay_x": replay_x.numpy(), "replay_labels": replay_labels.numpy(),
       "logits": logits_.detach().numpy(), "replay_termination_logits": replay_aux["termination_logits"].detach().numpy()}
for k, v in m.state_dict().items():
    if not k.endswith("attn.mask"):
        out["sd." + k] = v.detach().numpy()
for k, p in m.named_parameters():
    if p.grad is not None:
        out["grad." + k] = p.grad.detach().numpy()
parts = {"next": float(next_loss_), "offsets": {int(k): float(v) for k, v in off_losses.items()},
         "termination": float(term), "replay": float(replay_loss), "total": float(total)}
out["meta"] = np.array(json.dumps(dict(ctor=CTOR, parts=parts, offset_weights={str(k): v for k, v in OW.items()},
                                       termination_loss_weight=TW, replay_loss_weight=RW, replay_class_weights=REPLAY_CW,
                                       attention_window=None, torch=torch.__version__)))
np.savez_compressed(os.path.join(HERE, "replay_term.npz"), **out)
print(parts, os.path.getsize(os.path.join(HERE, "replay_term.npz")) // 1024, "KiB")
