"""Drop-in for the reference's ``src/codonlm/training/objectives.py`` (:6-105) on the CUDA kernels.

The integer parts (offset validity, termination labels) are scans over the target ids; the losses are
the fused cross-entropy kernel with a target shift and a validity test, so nothing of shape
(B, T, V) is gathered or copied.  ``training_loss`` is the loss composition of the trainer's ``fwd()``
(src/codonlm/training/loop.py:1067-1143) without its host syncs.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import functional as Fn
from . import ops

PAD_ID = 0
DEFAULT_BOUNDARY_IDS = (2, 3)  # <EOS_CDS>, <SEP>


def offset_target_mask(yb: torch.Tensor, offset: int, boundary_ids=DEFAULT_BOUNDARY_IDS) -> torch.Tensor:
    """bool (B, T-offset+1): target yb[:, t+offset-1] is not PAD and no boundary id lies in
    yb[:, t : t+offset-1]  (objectives.py:6-23)."""
    if offset < 1:
        raise ValueError("offset must be >= 1")
    B, T = yb.shape
    if offset > T:
        return torch.zeros((B, 0), dtype=torch.bool, device=yb.device)
    n = T - (offset - 1)
    yb = yb.contiguous()
    valid = yb[:, offset - 1:] != PAD_ID
    if offset > 1 and len(tuple(boundary_ids)) > 0:
        nb = ops.next_in_set(yb, tuple(boundary_ids))
        t = torch.arange(n, device=yb.device, dtype=torch.int32).unsqueeze(0)
        valid = valid & (nb[:, :n] >= t + (offset - 1))
    return valid


def multi_offset_lm_loss(logits, yb: torch.Tensor, offset_weights: Dict[int, float], label_smoothing: float = 0.0,
                         loss_weights: Optional[torch.Tensor] = None, boundary_ids=DEFAULT_BOUNDARY_IDS,
                         sync: bool = True):
    """(total, {offset: loss}) as objectives.py:26-60.  With sync=True offsets that have no valid target are
    left out of the dict exactly like the reference (one host sync for all offsets); with sync=False they
    contribute 0 and stay in the dict (no host sync: the training step uses this)."""
    yb = yb.contiguous()
    B, T = yb.shape
    nb = ops.next_in_set(yb, tuple(boundary_ids)) if len(tuple(boundary_ids)) > 0 else None
    losses, counts = {}, {}
    total = torch.zeros((), dtype=torch.float32, device=yb.device)
    for offset, weight in offset_weights.items():
        if weight == 0.0 or offset <= 1 or offset > T:
            continue
        if isinstance(logits, dict):
            if offset not in logits:
                continue
            lg = logits[offset]
        else:
            lg = logits
        lg2 = lg.reshape(B * T, lg.shape[-1])
        if lg2.dtype != torch.float32 or not lg2.is_contiguous():
            lg2 = lg2.float().contiguous()
        loss, sums = Fn.CrossEntropyFn.apply(lg2, yb, nb, loss_weights, B, T, offset - 1, float(label_smoothing),
                                             PAD_ID, True, getattr(lg, "_cgpt_grad_form", 0))
        losses[offset] = loss
        counts[offset] = sums[1]
        total = total + float(weight) * loss
    if sync and losses:
        kept = torch.stack([counts[o] for o in losses]).tolist()
        losses = {o: l for (o, l), c in zip(losses.items(), kept) if c > 0}
    return total, losses


def termination_distance_bucket_labels(yb: torch.Tensor, stop_ids: Sequence[int],
                                       bucket_edges: Sequence[int] = (0, 3, 10, 30),
                                       ignore_index: int = -100) -> torch.Tensor:
    """objectives.py:63-91 as one reverse scan + one elementwise kernel."""
    if not stop_ids:
        raise ValueError("stop_ids must not be empty")
    if tuple(bucket_edges) != tuple(sorted(bucket_edges)):
        raise ValueError("bucket_edges must be sorted")
    yb = yb.contiguous()
    nxt = ops.next_in_set(yb, tuple(stop_ids))
    return ops.termination_labels(yb, nxt, tuple(bucket_edges), ignore_index)


def termination_aux_loss(termination_logits: torch.Tensor, labels: torch.Tensor,
                         class_weights: Optional[torch.Tensor] = None, ignore_index: int = -100) -> torch.Tensor:
    """objectives.py:94-105."""
    B, T = labels.shape
    lg2 = termination_logits.reshape(B * T, termination_logits.shape[-1])
    if lg2.dtype != torch.float32 or not lg2.is_contiguous():
        lg2 = lg2.float().contiguous()
    loss, _ = Fn.CrossEntropyFn.apply(lg2, labels.contiguous(), None, class_weights, B, T, 0, 0.0, int(ignore_index),
                                      False)
    return loss


def training_loss(model, xb, yb, offset_weights: Optional[Dict[int, float]] = None,
                  termination_loss_weight: float = 0.0, termination_stop_ids=(2,),
                  termination_bucket_edges=(0, 3, 10, 30), termination_class_weights=None,
                  attention_window=None, shape_embeddings=None, replay=None, replay_loss_weight: float = 0.1,
                  replay_class_weights=None, replay_shape_embeddings=None):
    """total = next + sum_o w_o·loss_o + termination_loss_weight·term + replay_loss_weight·replay
    (the trainer's fwd(), loop.py:1067-1143).  `replay = (replay_x, replay_labels)` is the generated-state replay batch
    of this micro-batch (loop.py:1113-1141): a second forward whose termination logits are scored against sparse labels
    (-100 = unsupervised).  Returns (total, parts, logits); no host synchronisation."""
    need_aux = bool(offset_weights) or bool(termination_loss_weight)
    if need_aux:
        logits, next_loss, aux = model(xb, yb, return_aux=True, attention_window=attention_window,
                                       shape_embeddings=shape_embeddings)
    else:
        logits, next_loss = model(xb, yb, attention_window=attention_window, shape_embeddings=shape_embeddings)
        aux = {}
    yb = yb.to(logits.device)
    total = next_loss
    parts = {"next": next_loss}
    if offset_weights:
        off_total, off_losses = multi_offset_lm_loss(aux.get("offset_logits", logits), yb, offset_weights,
                                                     label_smoothing=model.label_smoothing,
                                                     loss_weights=model.class_weights(), sync=False)
        total = total + off_total
        parts["offsets"] = off_losses
    if termination_loss_weight:
        term_logits = aux.get("termination_logits")
        if term_logits is None:
            raise RuntimeError("termination_loss_enabled=true but model returned no termination logits")
        labels = termination_distance_bucket_labels(yb, termination_stop_ids, termination_bucket_edges)
        tl = termination_aux_loss(term_logits, labels, termination_class_weights)
        total = total + termination_loss_weight * tl
        parts["termination"] = tl
    if replay is not None:
        replay_x, replay_labels = replay
        _, _, replay_aux = model(replay_x, return_aux=True, shape_embeddings=replay_shape_embeddings)
        replay_logits = replay_aux.get("termination_logits")
        if replay_logits is None:
            raise RuntimeError("replay_loss_enabled=true but model returned no termination logits")
        rl = termination_aux_loss(replay_logits, replay_labels.to(replay_logits.device), replay_class_weights)
        total = total + replay_loss_weight * rl
        parts["replay"] = rl
    return total, parts, logits
