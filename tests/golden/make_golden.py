"""Generate golden vectors by running the UNMODIFIED reference implementation.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

For each case it builds the reference ``TinyGPT`` (src/codonlm/model_tiny_gpt.py:155) with
``torch.manual_seed``, perturbs a few parameters so that every branch is numerically
visible, runs forward / the trainer's loss composition (src/codonlm/training/loop.py:1067-1143
via the reference's own ``multi_offset_lm_loss`` / ``termination_*`` functions) / backward in
fp32 on CPU, and stores weights, inputs, outputs and gradients in ``tests/golden/<case>.npz``.
The committed ``.npz`` files are what the oracle and the CUDA path are checked against.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CGPT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from src.codonlm.model_tiny_gpt import TinyGPT  # noqa: E402
from src.codonlm.training import objectives as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.codon_gpt_oracle import synthetic_batch  # noqa: E402  (token generator only)

CASES = {
    # name: (ctor kwargs, B, T, extras)
    "gelu_abs_sep": dict(
        ctor=dict(vocab_size=68, block_size=48, n_layer=2, n_head=1, n_embd=32, dropout=0.0,
                  label_smoothing=0.05, sep_id=3, use_sdpa=True),
        B=3, T=40, emb_scale=0.02),
    "gelu_default_init": dict(
        ctor=dict(vocab_size=68, block_size=32, n_layer=2, n_head=1, n_embd=32, dropout=0.0,
                  label_smoothing=0.0, sep_id=3, use_sdpa=False),
        B=2, T=32, emb_scale=1.0),
    "swiglu_rope_causal": dict(
        ctor=dict(vocab_size=68, block_size=64, n_layer=2, n_head=2, n_embd=64, dropout=0.0,
                  label_smoothing=0.05, sep_id=None, use_sdpa=True, use_swiglu=True, use_rope=True),
        B=2, T=64, emb_scale=0.02),
    "gqa_untied_weighted": dict(
        ctor=dict(vocab_size=69, block_size=40, n_layer=1, n_head=4, n_kv_head=2, n_embd=128, dropout=0.0,
                  label_smoothing=0.1, sep_id=3, tie_embeddings=False, use_sdpa=True,
                  loss_weights=[1.0, 1.0, 3.0] + [1.0] * 66),
        B=2, T=33, emb_scale=0.02),
    "heads_offsets_term": dict(
        ctor=dict(vocab_size=68, block_size=64, n_layer=1, n_head=1, n_embd=64, dropout=0.0,
                  label_smoothing=0.05, sep_id=3, use_sdpa=True, termination_aux=True,
                  multi_offset_targets=[2, 4, 8]),
        B=3, T=64, emb_scale=0.02,
        offset_weights={2: 0.5, 4: 0.25, 8: 0.125}, termination_loss_weight=0.3),
    "shape_guidance": dict(
        ctor=dict(vocab_size=68, block_size=48, n_layer=1, n_head=2, n_embd=64, dropout=0.0,
                  label_smoothing=0.05, sep_id=3, use_sdpa=True, use_shape_guidance=True),
        B=2, T=40, emb_scale=0.02, shape_guidance=True),
    "window5": dict(
        ctor=dict(vocab_size=68, block_size=32, n_layer=1, n_head=1, n_embd=32, dropout=0.0,
                  label_smoothing=0.0, sep_id=3, use_sdpa=True),
        B=2, T=32, emb_scale=0.02, attention_window=5),
}


def state_sha256(model):
    import hashlib
    h = hashlib.sha256()
    for k, v in model.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build(case):
    spec = CASES[case]
    torch.manual_seed(1337)
    m = TinyGPT(**spec["ctor"])
    spec["init_sha256"] = state_sha256(m)  # weights straight after construction under manual_seed(1337)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        m.tok_emb.weight.mul_(spec["emb_scale"])
        if m.pos_emb is not None:
            m.pos_emb.weight.mul_(spec["emb_scale"])
        for name, p in m.named_parameters():
            if ".ln" in name or name.startswith("ln_f"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if name.startswith("offset_projs"):
                p.add_(0.02 * torch.randn(p.shape, generator=g))
            if name.startswith("shape_proj"):  # zero-initialised in the reference (:228-229): make it visible
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    m.eval()
    return m, spec


def main():
    for case in CASES:
        m, spec = build(case)
        V = spec["ctor"]["vocab_size"]
        idx, tgt = synthetic_batch(spec["B"], spec["T"], seed=11, realistic=True, vocab_size=V)
        win = spec.get("attention_window")
        out = {}
        shapes = None
        if spec.get("shape_guidance"):  # stands in for the DNA-shape encoder's output (B, T, 3), loop.py:1070-1077
            shapes = torch.randn((spec["B"], spec["T"], 3), generator=torch.Generator().manual_seed(5)).requires_grad_(True)
        logits, loss, aux = m(idx, tgt, return_aux=True, attention_window=win, shape_embeddings=shapes)
        total = loss
        parts = {"next": float(loss.detach())}
        ow = spec.get("offset_weights")
        if ow:
            lw = None if bool(torch.all(m.loss_weights == 1.0)) else m.loss_weights
            off_total, off_losses = O.multi_offset_lm_loss(
                aux["offset_logits"], tgt, ow, label_smoothing=spec["ctor"]["label_smoothing"], loss_weights=lw)
            total = total + off_total
            parts["offsets"] = {int(k): float(v) for k, v in off_losses.items()}
            for o, lg in aux["offset_logits"].items():
                out[f"offset_logits.{o}"] = lg.detach().numpy()
                out[f"offset_mask.{o}"] = O.offset_target_mask(tgt, o).numpy()
        tw = spec.get("termination_loss_weight", 0.0)
        if tw:
            labels = O.termination_distance_bucket_labels(tgt, stop_ids=(2,), bucket_edges=(0, 3, 10, 30))
            tl = O.termination_aux_loss(aux["termination_logits"], labels)
            total = total + tw * tl
            parts["termination"] = float(tl)
            out["termination_labels"] = labels.numpy()
        if "termination_logits" in aux:
            out["termination_logits"] = aux["termination_logits"].detach().numpy()
        total.backward()
        parts["total"] = float(total)
        mask = m.build_attention_mask(idx, win)
        out["attn_mask"] = (mask.numpy() if mask is not None else np.zeros((0,), dtype=bool))
        hidden = [h.detach().numpy() for _, h in m.iter_hidden_states(idx, attention_window=win,
                                                                      shape_embeddings=shapes)]
        if shapes is not None:
            out["shape_embeddings"] = shapes.detach().numpy()
            out["grad_shape_embeddings"] = shapes.grad.detach().numpy()
        out["hidden_final"] = hidden[-1]
        out["hidden_0"] = hidden[0]
        out["idx"] = idx.numpy()
        out["targets"] = tgt.numpy()
        out["logits"] = logits.detach().numpy()
        out["argmax"] = logits.detach().argmax(-1).numpy()
        for k, v in m.state_dict().items():
            if k.endswith("attn.mask"):
                continue  # tril(ones) buffer: reconstructible, 4*block^2 bytes per layer
            out["sd." + k] = v.detach().numpy()
        for k, p in m.named_parameters():
            if p.grad is not None:
                out["grad." + k] = p.grad.detach().numpy()
        meta = dict(ctor=spec["ctor"], parts=parts, attention_window=win,
                    offset_weights={str(k): v for k, v in (ow or {}).items()},
                    termination_loss_weight=tw, torch=torch.__version__, init_sha256=spec["init_sha256"])
        out["meta"] = np.array(json.dumps(meta))
        path = os.path.join(HERE, f"{case}.npz")
        np.savez_compressed(path, **out)
        print(case, {k: round(v, 6) if isinstance(v, float) else v for k, v in parts.items()},
              os.path.getsize(path) // 1024, "KiB")

    # integer-function vectors at larger, random shapes (reference objectives.py run directly)
    rng = np.random.default_rng(5)
    yb = rng.integers(0, 12, size=(6, 257), dtype=np.int64)
    yb[:, 200:] *= (rng.random((6, 57)) < 0.5)
    ints = {"yb": yb}
    t = torch.from_numpy(yb)
    for o in (1, 2, 3, 4, 8, 16, 32):
        ints[f"offset_mask.{o}"] = O.offset_target_mask(t, o).numpy()
    for name, (stops, edges) in {"a": ((2,), (0, 3, 10, 30)), "b": ((2, 3), (0, 1, 3)), "c": ((2,), ())}.items():
        ints[f"term.{name}"] = O.termination_distance_bucket_labels(t, stop_ids=stops, bucket_edges=edges).numpy()
    np.savez_compressed(os.path.join(HERE, "integer_kats.npz"), **ints)
    print("integer_kats ok")


if __name__ == "__main__":
    main()
