"""CPU test of the cached sampling loop's control flow (codonlm_b200/generate.py) with a stand-in model: it must pick
the same tokens as the reference's loop (scripts/query_model.py:186-213: full forward of the cropped context for every
token, temperature, top-k over probabilities, multinomial, eos stop), including past block_size where the context is
cropped and the cache has to be rebuilt."""
import torch

from codonlm_b200.generate import generate


class FakeState:
    def __init__(self, ctx, max_len):
        self.ctx = list(ctx)
        self.max_len = max_len

    @property
    def length(self):
        return len(self.ctx)


class FakeModel:
    """logits are a deterministic function of the WHOLE visible context (and of its absolute positions), so any
    mismatch between the cached path and the full re-run shows up as different tokens."""
    vocab = 11

    def __init__(self, block_size):
        self.block_size = block_size
        self.prefills = 0
        self.steps = 0

    def full_logits(self, ctx):
        g = torch.Generator().manual_seed(sum((i + 1) * (t + 3) for i, t in enumerate(ctx)) % (2 ** 31))
        return torch.randn(self.vocab, generator=g) * 3.0

    def prefill(self, x, max_len=None):
        self.prefills += 1
        ctx = x[0].tolist()
        return self.full_logits(ctx)[None], FakeState(ctx, max_len or self.block_size)

    def decode_step(self, tokens, state):
        self.steps += 1
        assert state.length < state.max_len
        state.ctx.append(int(tokens[0]))
        return self.full_logits(state.ctx)[None]


def reference_loop(model, ctx_ids, max_new, temperature, topk, eos_idx, seed):
    torch.manual_seed(seed)
    ids = list(ctx_ids)
    for _ in range(max_new):
        ctx = ids[-model.block_size:]
        logits = model.full_logits(ctx)
        if temperature != 1.0:
            logits = logits / max(1e-6, float(temperature))
        probs = torch.softmax(logits, dim=-1)
        if topk and topk > 0:
            vals, idxs = torch.topk(probs, k=min(topk, probs.numel()))
            next_id = idxs[torch.multinomial(vals, 1).item()].item()
        else:
            next_id = torch.multinomial(probs, 1).item()
        ids.append(next_id)
        if len(ids) > model.block_size:
            ids = ids[-model.block_size:]
        if eos_idx is not None and next_id == eos_idx:
            break
    return ids


def test_cached_loop_equals_reference_loop():
    for (block, n_ctx, max_new, temp, topk, eos) in [(16, 5, 8, 1.0, 0, None), (16, 5, 30, 0.7, 3, None),
                                                     (12, 12, 6, 1.0, 2, None), (32, 4, 25, 1.3, 0, 7)]:
        ctx = [1] + [4 + (3 * i) % 7 for i in range(n_ctx - 1)]
        want = reference_loop(FakeModel(block), ctx, max_new, temp, topk, eos, seed=5)
        model = FakeModel(block)
        torch.manual_seed(5)
        got = generate(model, "cpu", ctx, max_new, temperature=temp, topk=topk, eos_idx=eos)
        assert got == want, (block, n_ctx, max_new, temp, topk, eos)
        if n_ctx + max_new <= block and eos is None:
            assert model.prefills == 1 and model.steps == max_new - 1  # one prompt pass, then one position per token
