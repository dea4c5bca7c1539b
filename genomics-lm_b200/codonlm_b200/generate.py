"""Cached sampling loop: the reference's `generate` (scripts/query_model.py:186-213; src/codonlm/generate.py uses the
same next-token call, :14-27) with the context kept in a K/V cache instead of being re-run for every new token.

Same rule per step — logits / temperature, softmax, optional top-k over the probabilities, `torch.multinomial`, stop at
`eos_idx`, context cropped to the last `block_size` tokens — so a caller can swap it in for the reference function.
While the context is shorter than `block_size` a step costs one token's worth of work (TinyGPT.decode_step); once it
has to be cropped the positions of all cached tokens shift (absolute position embeddings), so each step re-prefills the
cropped context, which is exactly what the reference does on every step."""
from __future__ import annotations

from typing import List, Optional

import torch


@torch.no_grad()
def generate(model, device, ctx_ids: List[int], max_new: int, temperature: float = 1.0, topk: int = 0,
             eos_idx: Optional[int] = None) -> List[int]:
    ids = list(ctx_ids)
    max_t = getattr(model, "block_size", None)
    state = None
    pending = None  # the last sampled token: appended to the cache only when another token is asked for
    for _ in range(int(max_new)):
        if state is not None and pending is not None and state.length < state.max_len:
            logits = model.decode_step(torch.tensor([pending], device=device), state)
        else:  # first token, or the context was cropped (every cached position shifted): one pass over the context
            ctx = ids[-max_t:] if max_t is not None else ids
            x = torch.tensor(ctx, dtype=torch.long, device=device).unsqueeze(0)
            logits, state = model.prefill(x, max_len=max_t)
        row = logits[0]
        if temperature != 1.0:
            row = row / max(1e-6, float(temperature))
        probs = torch.softmax(row, dim=-1)
        if topk and topk > 0:
            vals, idxs = torch.topk(probs, k=min(topk, probs.numel()))
            next_id = int(idxs[torch.multinomial(vals, 1).item()].item())
        else:
            next_id = int(torch.multinomial(probs, 1).item())
        ids.append(next_id)
        pending = next_id
        if max_t is not None and len(ids) > max_t:
            ids = ids[-max_t:]
            state = None
        if eos_idx is not None and next_id == eos_idx:
            break
    return ids


@torch.no_grad()
def generate_batch(model, prompts: torch.Tensor, max_new: int, temperature: float = 1.0, topk: int = 0) -> torch.Tensor:
    """Batched variant for equal-length prompts (B, T0): one decode_step per new position for the whole batch.
    Returns (B, T0 + n) token ids, n = min(max_new, block_size - T0)."""
    dev = model.tok_emb.weight.device
    out = prompts.to(dev).long()
    logits, state = model.prefill(out)
    n = min(int(max_new), state.max_len - out.shape[1])
    for step in range(n):
        rows = logits if temperature == 1.0 else logits / max(1e-6, float(temperature))
        probs = torch.softmax(rows, dim=-1)
        if topk and topk > 0:
            vals, idxs = torch.topk(probs, k=min(topk, probs.shape[-1]), dim=-1)
            nxt = idxs.gather(-1, torch.multinomial(vals, 1))
        else:
            nxt = torch.multinomial(probs, 1)
        out = torch.cat([out, nxt], dim=1)
        if step + 1 < n:
            logits = model.decode_step(nxt.view(-1), state)
    return out
