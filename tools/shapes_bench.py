"""Throughput of the BASELINE.json configurations other than the headline one (which bench.py measures):

  C2  stage2.5 6L4H d256 RoPE+SwiGLU training step, seq 512
  C4  bench_b8_gqa4 (10L8H kv4 d384): training at batch 8 and a saturated batch, batched next-codon inference
  C5  long-context causal attention forward+backward, seq 4096, 8 heads (hd 48 and 64)

One JSON line per measurement (CUDA events, median of `--reps`, warm-up first, synthetic tokens).  These are
parity-test configurations, not bench lines; the numbers go to profiles/.  Usage: python tools/shapes_bench.py [c2 c4 c5]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
from bench import synthetic_tokens  # noqa: E402
from codonlm_b200 import TinyGPT, ops  # noqa: E402
from codonlm_b200.trainer import TrainStep  # noqa: E402

DEV = "cuda"
REPS = int(os.environ.get("REPS", "7"))


def timed(fn, reps=REPS, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def fwd_flops_per_token(L, d, T, V, hidden_mult, kv_frac=1.0):
    """SURVEY §8 convention: causal attention 2dT per layer; projections q,proj d² each, k,v kv_frac·d²."""
    per_layer = 2 * d * d * (2 + 2 * kv_frac) + 2 * d * T + 2 * hidden_mult * d * d
    return L * per_layer + 2 * d * V


def train_case(tag, ctor, B, T, flops_fwd):
    torch.manual_seed(1337)
    model = TinyGPT(**ctor)
    with torch.no_grad():
        model.tok_emb.weight.mul_(0.02)
    model = model.to(DEV).train()
    step = TrainStep(model, lr=3e-4, weight_decay=0.05)
    x, y = (t.to(DEV) for t in synthetic_tokens(B, T, seed=7))
    for _ in range(3):
        step.step(x, y)
    mode = "eager"
    try:
        step.capture(B, T)
        mode = "cuda graph"
    except Exception as exc:  # capture is an optimisation only
        print(f"[shapes] capture failed: {exc}", file=sys.stderr)
    ms = timed(lambda: step.step(x, y))
    toks = B * T / ms * 1e3
    emit(config=tag, what="train step fwd+bwd+AdamW", batch=B, seq=T, ms=ms, tokens_per_s=toks, execution=mode,
         model_tflops=toks * 3 * flops_fwd / 1e12, loss=float(step.step(x, y).item()))
    del step, model
    torch.cuda.empty_cache()


def c2():
    ctor = dict(vocab_size=68, block_size=512, n_layer=6, n_head=4, n_embd=256, dropout=0.0, label_smoothing=0.05,
                use_sdpa=True, use_rope=True, use_swiglu=True)
    f = fwd_flops_per_token(6, 256, 512, 68, 2 * 3 * 682 / 256)  # three d x 682 matrices
    for B in (8, 64, 256):
        train_case("C2 stage2.5 6L4H d256 RoPE+SwiGLU", ctor, B, 512, f)


def c4():
    ctor = dict(vocab_size=68, block_size=512, n_layer=10, n_head=8, n_kv_head=4, n_embd=384, dropout=0.0,
                label_smoothing=0.05, use_sdpa=True)
    f = fwd_flops_per_token(10, 384, 512, 68, 16, kv_frac=0.5)
    for B in (8, 128):
        train_case("C4 bench_b8_gqa4 10L8H kv4 d384", ctor, B, 512, f)
    torch.manual_seed(1337)
    model = TinyGPT(**ctor)
    with torch.no_grad():
        model.tok_emb.weight.mul_(0.02)
    model = model.to(DEV).eval()
    for B in (1, 8, 64, 512):
        x = synthetic_tokens(B, 512, seed=11)[0].to(DEV)
        with torch.no_grad():
            ms_full = timed(lambda: model(x)[0][:, -1].argmax(-1))
            ms_last = timed(lambda: model.next_token_logits(x).argmax(-1))
            same = bool(torch.equal(model(x)[0][:, -1].argmax(-1), model.next_token_logits(x).argmax(-1)))
        emit(config="C4 bench_b8_gqa4 10L8H kv4 d384", what="batched next-codon inference (last-position logits + argmax)",
             batch=B, seq=512, ms_forward_full_logits=ms_full, ms_next_token_logits=ms_last,
             contexts_per_s=B / ms_last * 1e3, context_tokens_per_s=B * 512 / ms_last * 1e3,
             model_tflops=B * 512 / ms_last * 1e3 * f / 1e12, argmax_equal_to_full_forward=same)


def decode():
    """C4 model: cached decode (prefill 256 tokens, then 128 decode steps) against the reference's procedure of one
    full forward per generated token (timed on this repo's own forward at the average context length)."""
    ctor = dict(vocab_size=68, block_size=512, n_layer=10, n_head=8, n_kv_head=4, n_embd=384, dropout=0.0,
                label_smoothing=0.05, use_sdpa=True)
    torch.manual_seed(1337)
    model = TinyGPT(**ctor)
    with torch.no_grad():
        model.tok_emb.weight.mul_(0.02)
    model = model.to(DEV).eval()
    T0, n_new = 256, 128
    for B in (1, 8, 64):
        seq = synthetic_tokens(B, T0 + n_new, seed=3)[0].to(DEV)

        with torch.no_grad():
            _, st = model.prefill(seq[:, :T0])
            for s in range(3):  # eager warm-up step, graph capture, first replay
                model.decode_step(seq[:, T0 + s], st)
            ts = []
            for rep in range(3):
                _, st = model.prefill(seq[:, :T0], state=st)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for s in range(n_new):
                    model.decode_step(seq[:, T0 + s], st)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / n_new)
            ms_step = sorted(ts)[1]
            ms_full = timed(lambda: model.next_token_logits(seq[:, :T0 + n_new // 2]), reps=3, warmup=1)
        emit(config="C4 bench_b8_gqa4 10L8H kv4 d384", what="KV-cache decode vs one full forward per token", batch=B,
             prompt=T0, new_tokens=n_new, ms_per_decode_step=ms_step, ms_per_full_forward_at_mean_context=ms_full,
             tokens_per_s_cached=B / ms_step * 1e3, tokens_per_s_full_forward=B / ms_full * 1e3,
             speedup=ms_full / ms_step)


def c5():
    B, T, H = 8, 4096, 8
    for hd in (48, 64):
        for mask in ("causal", "segment-causal"):
            qkv = torch.randn(B * T, 3 * H * hd, device=DEV).to(torch.bfloat16)
            ss = None
            if mask != "causal":
                idx = synthetic_tokens(B, T, seed=5)[0].to(DEV)
                ss = ops.segment_starts(idx, 3)
            out, lse = ops.attn_fwd(qkv, ss, B, T, H, H, hd)
            dout = torch.randn_like(out)
            ms_f = timed(lambda: ops.attn_fwd(qkv, ss, B, T, H, H, hd))
            ms_b = timed(lambda: ops.attn_bwd(qkv, ss, out, dout, lse, B, T, H, H, hd))
            # algorithmic pairs: (i, j) visible under the mask
            if ss is None:
                pairs = B * H * T * (T + 1) // 2
            else:
                pos = torch.arange(T, device=DEV)[None, :]
                pairs = int((pos - ss.long() + 1).sum().item()) * H
            emit(config="C5 long-context attention 8H seq 4096", head_dim=hd, mask=mask, batch=B, us_fwd=ms_f * 1e3,
                 us_bwd=ms_b * 1e3, visible_pairs=pairs, tflops_fwd=4 * hd * pairs / ms_f / 1e9,
                 tflops_bwd=10 * hd * pairs / ms_b / 1e9, tokens_per_s_fwd_bwd=B * T / (ms_f + ms_b) * 1e3)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c4", "c5"]
    for w in which:
        {"c2": c2, "c4": c4, "c5": c5, "decode": decode}[w]()
