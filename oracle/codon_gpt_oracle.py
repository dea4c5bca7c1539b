"""CPU oracle for the codon-GPT step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product
(``genomics-lm_b200/``) never does: it has no CPU path at all.

This is a *restatement*, in functional plain-PyTorch fp32 (plus numpy for the
integer parts), of the algorithm in the reference's

  src/codonlm/model_tiny_gpt.py      (RotaryEmbedding :9-45, SwiGLU :47-57,
                                      CausalSelfAttention :59-132, Block :134-153,
                                      TinyGPT.forward :297-352,
                                      build_attention_mask :273-295,
                                      iter_hidden_states :368-389)
  src/codonlm/training/objectives.py (offset_target_mask :6-23,
                                      multi_offset_lm_loss :26-60,
                                      termination_distance_bucket_labels :63-91,
                                      termination_aux_loss :94-105)
  src/codonlm/training/loop.py       (loss composition in fwd() :1067-1143)

It works on a *state dict* (the reference's checkpoint key names)
instead of an nn.Module tree, so identical weights are shared with the reference
and with the CUDA implementation by construction.

Parity pin: ``tests/golden/*.npz`` are produced by running the UNMODIFIED
reference (``/root/reference``) through ``tests/golden/make_golden.py``;
``tests/test_oracle_golden.py`` checks this oracle against every one of them
(logits, losses, gradients, masks, labels), and re-states the reference's own
known-answer tables (mask truth table tests/test_models.py:29-51, offset mask
tests/test_long_range_codon_objectives.py:15-32, termination labels :72-92).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Sequence

import numpy as np
import torch

PAD_ID = 0          # objectives.py:3
BOUNDARY_IDS = (2, 3)  # <EOS_CDS>, <SEP>; objectives.py:4
LN_EPS = 1e-5       # nn.LayerNorm default used at model_tiny_gpt.py:137,139,216
ROPE_BASE = 10000.0  # model_tiny_gpt.py:10


# --------------------------------------------------------------------------
# configuration helpers
# --------------------------------------------------------------------------
def make_cfg(vocab_size, block_size, n_layer=3, n_head=4, n_embd=256, dropout=0.1,
             use_checkpoint=False, label_smoothing=0.0, sep_id=3, tie_embeddings=True,
             n_kv_head=None, use_sdpa=False, loss_weights=None, termination_aux=False,
             termination_n_classes=5, multi_offset_targets=None, use_swiglu=False,
             use_rope=False, use_shape_guidance=False) -> dict:
    """Same argument list as TinyGPT.__init__ (model_tiny_gpt.py:156-177)."""
    if n_kv_head is not None and not (0 < int(n_kv_head) <= n_head):
        n_kv_head = None  # :64, :189
    return dict(vocab_size=int(vocab_size), block_size=int(block_size), n_layer=int(n_layer),
                n_head=int(n_head), n_embd=int(n_embd), dropout=float(dropout),
                label_smoothing=float(label_smoothing), sep_id=sep_id,
                tie_embeddings=bool(tie_embeddings), n_kv_head=n_kv_head,
                termination_aux=bool(termination_aux),
                termination_n_classes=int(termination_n_classes),
                multi_offset_targets=sorted({int(t) for t in multi_offset_targets}) if multi_offset_targets else [],
                use_swiglu=bool(use_swiglu), use_rope=bool(use_rope),
                use_shape_guidance=bool(use_shape_guidance))


# --------------------------------------------------------------------------
# integer / boolean parts (numpy; bit-exact bar)
# --------------------------------------------------------------------------
def segment_ids(idx: np.ndarray, sep_id: Optional[int]) -> np.ndarray:
    """Inclusive running count of separator tokens (model_tiny_gpt.py:290)."""
    idx = np.asarray(idx)
    if sep_id is None:
        return np.zeros_like(idx, dtype=np.int64)
    return np.cumsum(idx == int(sep_id), axis=1, dtype=np.int64)


def attention_mask(idx: np.ndarray, sep_id: Optional[int], attention_window: Optional[int] = None):
    """bool (B,1,T,T) or None; model_tiny_gpt.py:273-295.

    allowed[b,i,j] = (j <= i) and (i-j < window) and seg[b,i] == seg[b,j].
    """
    idx = np.asarray(idx)
    if attention_window is not None and int(attention_window) < 1:
        raise ValueError("attention_window must be at least 1")
    if sep_id is None and attention_window is None:
        return None
    T = idx.shape[1]
    i = np.arange(T)[:, None]
    j = np.arange(T)[None, :]
    allowed = (i - j) >= 0
    if attention_window is not None:
        allowed = allowed & ((i - j) < int(attention_window))
    allowed = np.broadcast_to(allowed[None, None], (idx.shape[0], 1, T, T)).copy()
    if sep_id is not None:
        seg = segment_ids(idx, sep_id)
        allowed &= (seg[:, :, None] == seg[:, None, :])[:, None]
    return allowed


def offset_target_mask(yb: np.ndarray, offset: int, boundary_ids: Iterable[int] = BOUNDARY_IDS) -> np.ndarray:
    """Valid (row, t) for predicting yb[:, t+offset-1] from position t; objectives.py:6-23."""
    yb = np.asarray(yb)
    if offset < 1:
        raise ValueError("offset must be >= 1")
    B, T = yb.shape
    if offset > T:
        return np.zeros((B, 0), dtype=bool)
    n = T - (offset - 1)
    out = np.zeros((B, n), dtype=bool)
    bset = {int(b) for b in boundary_ids}
    for b in range(B):
        for t in range(n):
            ok = yb[b, t + offset - 1] != PAD_ID
            # an EOS/SEP strictly before the target position (within the hop) kills it
            for s in range(offset - 1):
                if int(yb[b, t + s]) in bset:
                    ok = False
                    break
            out[b, t] = ok
    return out


def termination_distance_bucket_labels(yb: np.ndarray, stop_ids: Sequence[int],
                                       bucket_edges: Sequence[int] = (0, 3, 10, 30),
                                       ignore_index: int = -100) -> np.ndarray:
    """objectives.py:63-91 as a right-to-left scan."""
    if not stop_ids:
        raise ValueError("stop_ids must not be empty")
    if tuple(bucket_edges) != tuple(sorted(bucket_edges)):
        raise ValueError("bucket_edges must be sorted")
    yb = np.asarray(yb)
    B, T = yb.shape
    stops = {int(s) for s in stop_ids}
    labels = np.full((B, T), ignore_index, dtype=np.int64)
    for b in range(B):
        nxt = T  # sentinel: no stop at/after this position
        for t in range(T - 1, -1, -1):
            if int(yb[b, t]) in stops:
                nxt = t
            if yb[b, t] == PAD_ID:
                continue
            if nxt == T:
                labels[b, t] = len(bucket_edges)
            else:
                d = nxt - t
                labels[b, t] = sum(1 for e in bucket_edges if d > e)
    return labels


# --------------------------------------------------------------------------
# floating-point building blocks (torch fp32, differentiable)
# --------------------------------------------------------------------------
def layer_norm(x, w, b, eps=LN_EPS):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x):
    """nn.GELU() default = exact erf form (model_tiny_gpt.py:145)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def silu(x):
    return x * torch.sigmoid(x)


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def rope_tables(T: int, head_dim: int, device=None):
    """cos/sin (T, head_dim), half-split layout: emb = cat(freqs, freqs); :15-25."""
    inv_freq = 1.0 / (ROPE_BASE ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    t = torch.arange(T, dtype=torch.float32)
    f = torch.outer(t, inv_freq)
    emb = torch.cat((f, f), dim=-1)
    return emb.cos().to(device), emb.sin().to(device)


def apply_rope(x, cos, sin):
    """x: (B,H,T,hd).  x*cos + cat(-x2, x1)*sin, pairing i <-> i+hd/2 (:35-45)."""
    h = x.shape[-1] // 2
    rot = torch.cat((-x[..., h:], x[..., :h]), dim=-1)
    return x * cos[None, None] + rot * sin[None, None]


def cross_entropy(logits, targets, ignore_index=PAD_ID, label_smoothing=0.0, weight=None):
    """F.cross_entropy(mean) semantics spelled out (model_tiny_gpt.py:343-349).

    loss = sum_{i: y_i != ignore} [ (1-eps) * w[y_i] * (-logp_i[y_i])
                                    + (eps / V) * sum_c w[c] * (-logp_i[c]) ]
           / sum_{i: y_i != ignore} w[y_i]
    An all-ignored batch gives 0/0 = NaN like torch.
    """
    logits = logits.float()
    V = logits.shape[-1]
    logp = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    keep = targets != ignore_index
    safe_t = torch.where(keep, targets, torch.zeros_like(targets))
    w = torch.ones(V, dtype=torch.float32, device=logits.device) if weight is None else weight.float()
    wt = w[safe_t] * keep
    nll = -(logp.gather(-1, safe_t[..., None])[..., 0]) * wt
    loss_sum = (1.0 - label_smoothing) * nll.sum()
    if label_smoothing > 0.0:
        smooth = -(logp * w).sum(dim=-1) * keep
        loss_sum = loss_sum + (label_smoothing / V) * smooth.sum()
    return loss_sum / wt.sum()


# --------------------------------------------------------------------------
# the model step
# --------------------------------------------------------------------------
def _attention(sd, pre, x, cfg, mask_bool, cos, sin):
    B, T, C = x.shape
    H = cfg["n_head"]
    Hk = cfg["n_kv_head"] or H
    if H % Hk != 0:
        raise ValueError("n_head must be divisible by n_kv_head for GQA")
    hd = C // H
    q = linear(x, sd[pre + "query.weight"], sd[pre + "query.bias"]).view(B, T, H, hd).transpose(1, 2)
    k = linear(x, sd[pre + "key.weight"], sd[pre + "key.bias"]).view(B, T, Hk, hd).transpose(1, 2)
    v = linear(x, sd[pre + "value.weight"], sd[pre + "value.bias"]).view(B, T, Hk, hd).transpose(1, 2)
    if Hk != H:  # :94-96, query head h reads kv head h // rep
        rep = H // Hk
        k = k.repeat_interleave(rep, dim=1)
        v = v.repeat_interleave(rep, dim=1)
    if cos is not None:
        q = apply_rope(q, cos, sin)
        k = apply_rope(k, cos, sin)
    att = (q @ k.transpose(-2, -1)) / math.sqrt(hd)
    att = att.masked_fill(~mask_bool, float("-inf"))
    att = torch.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).reshape(B, T, C)
    return linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"]), att


def _mlp(sd, pre, x, cfg):
    if cfg["use_swiglu"]:
        g = linear(x, sd[pre + "w_gate.weight"])
        u = linear(x, sd[pre + "w_up.weight"])
        return linear(silu(g) * u, sd[pre + "w_down.weight"])
    h = gelu_erf(linear(x, sd[pre + "0.weight"], sd[pre + "0.bias"]))
    return linear(h, sd[pre + "2.weight"], sd[pre + "2.bias"])


def forward(sd: Dict[str, torch.Tensor], cfg: dict, idx: torch.Tensor,
            targets: Optional[torch.Tensor] = None, attention_window: Optional[int] = None,
            want_hidden: bool = False, want_attn: bool = False,
            shape_embeddings: Optional[torch.Tensor] = None) -> dict:
    """Eval-mode (dropout off) forward of TinyGPT (model_tiny_gpt.py:297-352).

    Returns dict(logits, loss|None, termination_logits?, offset_logits?{o:..},
                 hidden?[x0, x1.., ln_f(x)], attn?[per layer (B,H,T,T)]).
    """
    B, T = idx.shape
    dev = idx.device
    x = sd["tok_emb.weight"][idx]
    if not cfg["use_rope"]:
        x = x + sd["pos_emb.weight"][:T][None]
        cos = sin = None
    else:
        cos, sin = rope_tables(T, cfg["n_embd"] // cfg["n_head"], device=dev)
    if shape_embeddings is not None and cfg.get("use_shape_guidance"):  # model_tiny_gpt.py:310-311
        x = x + linear(shape_embeddings, sd["shape_proj.weight"], sd["shape_proj.bias"])
    m = attention_mask(idx.cpu().numpy(), cfg["sep_id"], attention_window)
    if m is None:
        m = np.tril(np.ones((T, T), dtype=bool))[None, None]
    mask_bool = torch.from_numpy(np.ascontiguousarray(m)).to(dev)
    hidden = [x]
    attn = []
    for l in range(cfg["n_layer"]):
        p = f"blocks.{l}."
        a, att = _attention(sd, p + "attn.", layer_norm(x, sd[p + "ln1.weight"], sd[p + "ln1.bias"]),
                            cfg, mask_bool, cos, sin)
        x = x + a
        x = x + _mlp(sd, p + "mlp.", layer_norm(x, sd[p + "ln2.weight"], sd[p + "ln2.bias"]), cfg)
        hidden.append(x)
        if want_attn:
            attn.append(att)
    x = layer_norm(x, sd["ln_f.weight"], sd["ln_f.bias"])
    hidden.append(x)
    head_w = sd["tok_emb.weight"] if cfg["tie_embeddings"] else sd["head.weight"]
    out = {"logits": linear(x, head_w)}
    if cfg["termination_aux"]:
        out["termination_logits"] = linear(x, sd["termination_head.weight"], sd["termination_head.bias"])
    if cfg["multi_offset_targets"]:
        ol = {}
        for o in cfg["multi_offset_targets"]:
            q = f"offset_projs.{o}."
            h = gelu_erf(linear(x, sd[q + "0.weight"], sd[q + "0.bias"]))
            ol[o] = linear(linear(h, sd[q + "2.weight"], sd[q + "2.bias"]), head_w)
        out["offset_logits"] = ol
    out["loss"] = None
    if targets is not None:
        lw = sd.get("loss_weights")
        w = None if (lw is None or bool(torch.all(lw == 1.0))) else lw
        out["loss"] = cross_entropy(out["logits"].reshape(-1, out["logits"].shape[-1]), targets.reshape(-1),
                                    ignore_index=PAD_ID, label_smoothing=cfg["label_smoothing"], weight=w)
    if want_hidden:
        out["hidden"] = hidden
    if want_attn:
        out["attn"] = attn
    return out


def multi_offset_lm_loss(offset_logits: Dict[int, torch.Tensor], yb: torch.Tensor,
                         offset_weights: Dict[int, float], label_smoothing=0.0, loss_weights=None,
                         boundary_ids=BOUNDARY_IDS):
    """objectives.py:26-60 (dict-logits form).  Returns (total, {offset: loss})."""
    losses = {}
    total = torch.zeros((), dtype=torch.float32, device=yb.device)
    T = yb.shape[1]
    for o, wgt in offset_weights.items():
        if wgt == 0.0 or o <= 1 or o > T or o not in offset_logits:
            continue
        valid = torch.from_numpy(offset_target_mask(yb.cpu().numpy(), o, boundary_ids)).to(yb.device)
        if not bool(valid.any()):
            continue
        tgt = yb[:, o - 1:]
        pred = offset_logits[o][:, : tgt.shape[1], :]
        # invalid positions are dropped == treated as ignore_index
        tgt_masked = torch.where(valid, tgt, torch.full_like(tgt, PAD_ID))
        l = cross_entropy(pred.reshape(-1, pred.shape[-1]), tgt_masked.reshape(-1), ignore_index=PAD_ID,
                          label_smoothing=label_smoothing, weight=loss_weights)
        losses[o] = l
        total = total + float(wgt) * l
    return total, losses


def termination_aux_loss(term_logits, labels, class_weights=None, ignore_index=-100):
    """objectives.py:94-105."""
    return cross_entropy(term_logits.reshape(-1, term_logits.shape[-1]), labels.reshape(-1),
                         ignore_index=ignore_index, label_smoothing=0.0, weight=class_weights)


def training_loss(sd, cfg, idx, targets, offset_weights=None, termination_loss_weight=0.0,
                  termination_stop_ids=(2,), termination_bucket_edges=(0, 3, 10, 30),
                  termination_class_weights=None, attention_window=None, shape_embeddings=None,
                  replay=None, replay_loss_weight=0.1, replay_class_weights=None):
    """Loss composition of the trainer's fwd() (loop.py:1067-1143).

    total = next + sum_o w_o * loss_o + termination_loss_weight * term + replay_loss_weight * replay
    where replay (loop.py:1113-1141) is the termination-head loss of a SECOND forward over a batch of generated
    contexts `replay = (replay_x, replay_labels)` (labels -100 except the supervised generated states).
    Returns (total, parts dict, forward-output dict).
    """
    out = forward(sd, cfg, idx, targets, attention_window=attention_window, shape_embeddings=shape_embeddings)
    total = out["loss"]
    parts = {"next": out["loss"]}
    if offset_weights:
        lw = sd.get("loss_weights")
        w = None if (lw is None or bool(torch.all(lw == 1.0))) else lw
        off_total, off_losses = multi_offset_lm_loss(out.get("offset_logits", {}), targets, offset_weights,
                                                     label_smoothing=cfg["label_smoothing"], loss_weights=w)
        total = total + off_total
        parts["offsets"] = off_losses
    if termination_loss_weight and cfg["termination_aux"]:
        labels = torch.from_numpy(termination_distance_bucket_labels(
            targets.cpu().numpy(), termination_stop_ids, termination_bucket_edges)).to(targets.device)
        tl = termination_aux_loss(out["termination_logits"], labels, termination_class_weights)
        total = total + termination_loss_weight * tl
        parts["termination"] = tl
    if replay is not None:
        replay_x, replay_labels = replay
        rout = forward(sd, cfg, replay_x, None)  # model(replay_x, return_aux=True): no targets, no shape guidance here
        rl = termination_aux_loss(rout["termination_logits"], replay_labels, replay_class_weights)
        total = total + replay_loss_weight * rl
        parts["replay"] = rl
    return total, parts, out


def loss_and_grads(sd, cfg, idx, targets, **kw):
    """fp32 autograd through the oracle: returns (total, parts, out, {name: grad})."""
    leaves = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point and k != "loss_weights" and not k.endswith("attn.mask"):
            leaves[k] = v.detach().clone().requires_grad_(True)
        else:
            leaves[k] = v
    if cfg["tie_embeddings"] and "head.weight" in leaves:
        leaves["head.weight"] = leaves["tok_emb.weight"]
    total, parts, out = training_loss(leaves, cfg, idx, targets, **kw)
    total.backward()
    grads = {k: v.grad for k, v in leaves.items()
             if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None
             and not (cfg["tie_embeddings"] and k == "head.weight")}
    return total.detach(), parts, out, grads


# --------------------------------------------------------------------------
# weights / synthetic data
# --------------------------------------------------------------------------
def init_state_dict(cfg: dict, seed: int = 1337, emb_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random weights with the reference's default distributions (not its RNG stream).

    nn.Embedding ~ N(0,1); nn.Linear weight ~ U(+-1/sqrt(fan_in)) (kaiming_uniform a=sqrt5),
    bias ~ U(+-1/sqrt(fan_in)); LayerNorm (1,0); offset MLPs identity (model_tiny_gpt.py:240-245).
    ``emb_scale`` = 0.02 gives the "trained-scale" variant of SURVEY §8d.
    """
    g = torch.Generator().manual_seed(seed)
    V, d, L = cfg["vocab_size"], cfg["n_embd"], cfg["n_layer"]
    H = cfg["n_head"]
    Hk = cfg["n_kv_head"] or H
    kv = Hk * (d // H)

    def lin(out_f, in_f, bias=True, pre=""):
        bound = 1.0 / math.sqrt(in_f)
        r = {pre + "weight": (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound}
        if bias:
            r[pre + "bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return r

    sd = {"loss_weights": torch.ones(V)}
    sd["tok_emb.weight"] = torch.randn(V, d, generator=g) * emb_scale
    if not cfg["use_rope"]:
        sd["pos_emb.weight"] = torch.randn(cfg["block_size"], d, generator=g) * emb_scale
    for l in range(L):
        p = f"blocks.{l}."
        sd[p + "ln1.weight"] = torch.ones(d) + 0.1 * torch.randn(d, generator=g)
        sd[p + "ln1.bias"] = 0.1 * torch.randn(d, generator=g)
        sd.update(lin(kv, d, pre=p + "attn.key."))
        sd.update(lin(d, d, pre=p + "attn.query."))
        sd.update(lin(kv, d, pre=p + "attn.value."))
        sd.update(lin(d, d, pre=p + "attn.proj."))
        sd[p + "ln2.weight"] = torch.ones(d) + 0.1 * torch.randn(d, generator=g)
        sd[p + "ln2.bias"] = 0.1 * torch.randn(d, generator=g)
        if cfg["use_swiglu"]:
            hdim = int(8 * d // 3)
            sd.update(lin(hdim, d, bias=False, pre=p + "mlp.w_gate."))
            sd.update(lin(hdim, d, bias=False, pre=p + "mlp.w_up."))
            sd.update(lin(d, hdim, bias=False, pre=p + "mlp.w_down."))
        else:
            sd.update(lin(4 * d, d, pre=p + "mlp.0."))
            sd.update(lin(d, 4 * d, pre=p + "mlp.2."))
    sd["ln_f.weight"] = torch.ones(d) + 0.1 * torch.randn(d, generator=g)
    sd["ln_f.bias"] = 0.1 * torch.randn(d, generator=g)
    sd["head.weight"] = sd["tok_emb.weight"] if cfg["tie_embeddings"] else \
        torch.randn(V, d, generator=g) * (emb_scale if emb_scale != 1.0 else 1.0 / math.sqrt(d))
    if cfg["termination_aux"]:
        sd.update(lin(cfg["termination_n_classes"], d, pre="termination_head."))
    for o in cfg["multi_offset_targets"]:
        q = f"offset_projs.{o}."
        # identity init perturbed so that the offset branch is numerically distinguishable
        sd[q + "0.weight"] = torch.eye(d) + 0.02 * torch.randn(d, d, generator=g)
        sd[q + "0.bias"] = 0.02 * torch.randn(d, generator=g)
        sd[q + "2.weight"] = torch.eye(d) + 0.02 * torch.randn(d, d, generator=g)
        sd[q + "2.bias"] = 0.02 * torch.randn(d, generator=g)
    if cfg.get("use_shape_guidance"):
        # zero-initialised in the reference (model_tiny_gpt.py:228-229); perturbed so that the branch is visible.
        # Drawn last: the weights of every other configuration keep their values.
        sd["shape_proj.weight"] = 0.05 * torch.randn(d, 3, generator=g)
        sd["shape_proj.bias"] = 0.05 * torch.randn(d, generator=g)
    return sd


def synthetic_batch(B: int, T: int, seed: int = 1337, realistic: bool = False, vocab_size: int = 68,
                    pad_tail_frac: float = 0.25):
    """(idx, targets) int64 (B,T).  Plain: codons U{4..V-1}, targets = idx shifted left, last col PAD.

    realistic: BOS first, EOS(2)+SEP(3) every U{100..400} tokens, and a PAD tail on a
    fraction of rows."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(4, vocab_size, size=(B, T), dtype=np.int64)
    if realistic:
        for b in range(B):
            idx[b, 0] = 1
            t = int(rng.integers(100, 401)) if T > 128 else int(rng.integers(4, max(5, T // 2)))
            while t + 1 < T:
                idx[b, t] = 2
                idx[b, t + 1] = 3
                t += int(rng.integers(100, 401)) if T > 128 else int(rng.integers(4, max(5, T // 2)))
            if rng.random() < 0.5:
                n_pad = int(rng.integers(1, max(2, int(2 * pad_tail_frac * T))))
                idx[b, T - n_pad:] = 0
    tgt = np.zeros_like(idx)
    tgt[:, :-1] = idx[:, 1:]
    return torch.from_numpy(idx), torch.from_numpy(tgt)


# ---------------------------------------------------------------------------------------------------------
# token feed (reference: src/codonlm/data_loading.py, MmapPackedDataset.fetch_batch, dynamic format :297-315)
# ---------------------------------------------------------------------------------------------------------
def fetch_batch_dynamic(flat: np.ndarray, lengths: np.ndarray, indices, pad_id: int = PAD_ID):
    """(xb, yb) int64 of one micro-batch from a flat token array + per-sequence lengths: rows are seq[:-1] / seq[1:]
    padded with PAD to max(len)-1 columns (data_loading.py:297-315; offsets as built at :222)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lengths[:-1])]).astype(np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    if indices.size == 0:
        return np.zeros((0, 0), np.int64), np.zeros((0, 0), np.int64)
    width = max(0, int(lengths[indices].max()) - 1)
    xb = np.full((indices.size, width), pad_id, dtype=np.int64)
    yb = np.full((indices.size, width), pad_id, dtype=np.int64)
    for r, s in enumerate(indices):
        seq = np.asarray(flat[offsets[s]: offsets[s] + lengths[s]], dtype=np.int64)
        usable = max(0, int(lengths[s]) - 1)
        if usable:
            xb[r, :usable] = seq[:-1]
            yb[r, :usable] = seq[1:]
    return xb, yb


# ---------------------------------------------------------------------------------------------------------
# NoProp variant (reference: model_tiny_gpt.py:391-459)
# ---------------------------------------------------------------------------------------------------------
def noprop_forward(sd: Dict[str, torch.Tensor], cfg: dict, idx: torch.Tensor,
                   target_embeddings: Optional[torch.Tensor] = None):
    """NoPropTinyGPT.forward (eval mode): (logits, [pred_y per block]).  Every block sees h + target_embeddings
    (:407-408), runs the usual pre-LN attention + GELU MLP (:412-413) and applies its denoise_head (:415)."""
    B, T = idx.shape
    h = sd["tok_emb.weight"][idx] + sd["pos_emb.weight"][:T][None]
    if cfg["sep_id"] is not None:
        m = attention_mask(idx.cpu().numpy(), cfg["sep_id"], None)
    else:
        m = np.tril(np.ones((T, T), dtype=bool))[None, None]
    mask_bool = torch.from_numpy(np.ascontiguousarray(m)).to(idx.device)
    sub = dict(cfg, use_swiglu=False)
    preds = []
    for l in range(cfg["n_layer"]):
        p = f"blocks.{l}."
        x = h + target_embeddings if target_embeddings is not None else h
        a, _ = _attention(sd, p + "attn.", layer_norm(x, sd[p + "ln1.weight"], sd[p + "ln1.bias"]), sub, mask_bool,
                          None, None)
        x = x + a
        x = x + _mlp(sd, p + "mlp.", layer_norm(x, sd[p + "ln2.weight"], sd[p + "ln2.bias"]), sub)
        preds.append(linear(x, sd[p + "denoise_head.weight"], sd[p + "denoise_head.bias"]))
        h = x
    h = layer_norm(h, sd["ln_f.weight"], sd["ln_f.bias"])
    return linear(h, sd["tok_emb.weight"]), preds
