"""Pins the oracle at the BASELINE.json shapes themselves (C1 'tiny 2L4H d128 seq 256 fp32 forward+loss on CPU', C2
'stage2.5 6L4H d256 RoPE+SwiGLU seq 512'), where full weight dumps would be megabytes: the weights come from the oracle's
own deterministic initialiser (so the test can rebuild them), are loaded STRICTLY into the unmodified reference model, and
only the reference's outputs are stored — loss, every 97th logit, the argmax map, hidden-state checksums.

    python tests/golden/make_baseline_shape_golden.py      (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.environ.get("CGPT_REFERENCE", "/root/reference"))
from src.codonlm.model_tiny_gpt import TinyGPT  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import codon_gpt_oracle as O  # noqa: E402  (weights + tokens only; the numbers below come from the reference)

CASES = {
    "C1": (dict(vocab_size=68, block_size=256, n_layer=2, n_head=4, n_embd=128, dropout=0.0, label_smoothing=0.05,
                use_sdpa=True), 8, 256),
    "C2": (dict(vocab_size=68, block_size=512, n_layer=6, n_head=4, n_embd=256, dropout=0.0, label_smoothing=0.05,
                use_sdpa=True, use_rope=True, use_swiglu=True), 2, 512),
    # C3: 12L8H d512, separate / multi-offset heads + termination head, seq 1024 (one sequence: forward + all losses)
    "C3": (dict(vocab_size=68, block_size=1024, n_layer=12, n_head=8, n_embd=512, dropout=0.0, label_smoothing=0.05,
                use_sdpa=True, termination_aux=True, multi_offset_targets=[2, 4, 8, 16, 32]), 1, 1024),
    # C4: bench_b8_gqa4 (10L8H, 4 kv heads, d384), seq 512
    "C4": (dict(vocab_size=68, block_size=512, n_layer=10, n_head=8, n_kv_head=4, n_embd=384, dropout=0.0,
                label_smoothing=0.05, use_sdpa=True), 2, 512),
}
from src.codonlm.training import objectives as RO  # noqa: E402
out = {}
for name, (ctor, B, T) in CASES.items():
    cfg = O.make_cfg(**ctor)
    sd = O.init_state_dict(cfg, seed=1337, emb_scale=0.02)
    m = TinyGPT(**ctor).eval()
    full = dict(sd)
    for l in range(ctor["n_layer"]):
        full[f"blocks.{l}.attn.mask"] = m.blocks[l].attn.mask
    missing = m.load_state_dict(full, strict=True)
    idx, tgt = O.synthetic_batch(B, T, seed=1337, realistic=True)
    with torch.no_grad():
        logits, loss, aux = m(idx, tgt, return_aux=True)
        hidden = m.forward_hidden(idx)
        if ctor.get("multi_offset_targets"):
            ow = {o: 0.2 for o in ctor["multi_offset_targets"]}
            off_total, off_losses = RO.multi_offset_lm_loss(aux["offset_logits"], tgt, ow,
                                                            label_smoothing=ctor["label_smoothing"], loss_weights=None)
            out[name + ".offset_losses"] = np.array([float(off_losses[o]) for o in ctor["multi_offset_targets"]])
            labels = RO.termination_distance_bucket_labels(tgt, stop_ids=(2,), bucket_edges=(0, 3, 10, 30))
            out[name + ".termination_loss"] = np.array(float(RO.termination_aux_loss(aux["termination_logits"], labels)))
            out[name + ".offset32_logit_samples"] = aux["offset_logits"][32].reshape(-1)[::97].numpy()
    # gradients of the trainer's total loss (loop.py:1067-1143): per-parameter norms and every 997th entry
    m.zero_grad()
    lg, ls, ax = m(idx, tgt, return_aux=True)
    total = ls
    if ctor.get("multi_offset_targets"):
        ot, _ = RO.multi_offset_lm_loss(ax["offset_logits"], tgt, {o: 0.2 for o in ctor["multi_offset_targets"]},
                                        label_smoothing=ctor["label_smoothing"], loss_weights=None)
        lb = RO.termination_distance_bucket_labels(tgt, stop_ids=(2,), bucket_edges=(0, 3, 10, 30))
        total = total + ot + 0.1 * RO.termination_aux_loss(ax["termination_logits"], lb)
    total.backward()
    names = [k for k, p in m.named_parameters() if p.grad is not None]
    out[name + ".grad_names"] = np.array(json.dumps(names))
    out[name + ".grad_norms"] = np.array([float(dict(m.named_parameters())[k].grad.norm()) for k in names])
    out[name + ".grad_samples"] = torch.cat([dict(m.named_parameters())[k].grad.reshape(-1)[::997] for k in names]).numpy()
    out[name + ".total_loss"] = np.array(float(total))
    flat = logits.reshape(-1)
    out[name + ".loss"] = np.array(float(loss), dtype=np.float64)
    out[name + ".logit_samples"] = flat[::97].numpy()
    out[name + ".argmax"] = logits.argmax(-1).numpy().astype(np.int8)
    out[name + ".logits_abs_mean"] = np.array(float(flat.abs().mean()), dtype=np.float64)
    out[name + ".hidden_row_norms"] = hidden.norm(dim=-1).numpy()
    out[name + ".meta"] = np.array(json.dumps(dict(ctor=ctor, B=B, T=T, torch=torch.__version__)))
    print(name, "loss", float(loss), "logits", tuple(logits.shape))
np.savez_compressed(os.path.join(HERE, "baseline_shapes.npz"), **out)
print("wrote baseline_shapes.npz", os.path.getsize(os.path.join(HERE, "baseline_shapes.npz")) // 1024, "KiB")
