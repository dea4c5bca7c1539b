#!/usr/bin/env python
"""Benchmark of the codon-GPT step (BASELINE.json metric: training codon tokens/sec).

  python bench.py --gpus 1 --steps K --warmup W                  # this repo's CUDA path, headline workload (C3)
  python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own CPU path, rank 0
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # data-parallel
  python bench.py --workload {c2,c4_train,c4_infer,c5_attn} ...   # the other BASELINE.json configurations

Headline workload (config.workload): BASELINE.json configs[2] = "12L8H d512 separate/multi-offset heads training
step, seq 1024" (the configuration the metric is quoted on; it fits one GPU): per-GPU micro-batch of 64 sequences x
1024 codons, offsets [2,4,8,16,32] + termination head, label smoothing 0.05, AdamW, on north_star's "synthetic
random-codon batches" (--tokens random; --tokens realistic is the segmented stream, credited for the attention work it
actually executes).  A step = forward + backward (+ gradient all-reduce) + optimiser update on one synthetic batch.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")  # the unmodified reference, vendored by __graft_entry__.build()

METRIC = "training codon tokens/sec"
UNIT = "tokens/s"
OFFSETS = [2, 4, 8, 16, 32]
TERM_W = 0.1

# BASELINE.json configs (SURVEY §8 tags).  kind: "train" = fwd+bwd+AdamW step; "infer" = batched next-codon inference;
# "attn" = attention forward+backward only.
WORKLOADS = {
    "c3": dict(kind="train", batch=64, seq=1024, offsets=OFFSETS,
               ctor=dict(n_layer=12, n_head=8, n_embd=512, termination_aux=True, multi_offset_targets=OFFSETS),
               name="C3 codon-GPT {L}L8H d512 (MHA hd64, GELU 2048) + offset heads [2, 4, 8, 16, 32] + termination head, "
                    "train step fwd+bwd+AdamW, seq {T}"),
    "c2": dict(kind="train", batch=64, seq=512, offsets=None,
               ctor=dict(n_layer=6, n_head=4, n_embd=256, use_rope=True, use_swiglu=True),
               name="C2 stage2.5 {L}L4H d256 RoPE+SwiGLU, train step fwd+bwd+AdamW, seq {T}"),
    "c4_train": dict(kind="train", batch=8, seq=512, offsets=None,
                     ctor=dict(n_layer=10, n_head=8, n_kv_head=4, n_embd=384),
                     name="C4 bench_b8_gqa4 {L}L8H kv4 d384 (hd48, GELU 1536), train step fwd+bwd+AdamW, seq {T}"),
    "c4_infer": dict(kind="infer", batch=512, seq=512, offsets=None,
                     ctor=dict(n_layer=10, n_head=8, n_kv_head=4, n_embd=384),
                     name="C4 bench_b8_gqa4 {L}L8H kv4 d384: batched next-codon inference (last-position logits + argmax), "
                          "contexts of {T} codons"),
    "c5_attn": dict(kind="attn", batch=8, seq=4096, offsets=None, ctor=dict(n_layer=1, n_head=8, n_embd=512),
                    name="C5 stage2_long_context: causal attention forward+backward, 8 heads hd64, seq {T}"),
}


def model_ctor(wl, n_layer=None, seq=None, dropout=0.0):
    c = dict(vocab_size=68, block_size=seq or wl["seq"], dropout=dropout, label_smoothing=0.05, sep_id=3, use_sdpa=True)
    c.update(wl["ctor"])
    if n_layer:
        c["n_layer"] = n_layer
    return c


def fwd_flops_per_token(ctor, T, visible_keys_per_token=None, V=68):
    """Algorithmic forward FLOPs per token.  Attention is credited for the (query, key) pairs the mask makes visible:
    4·d FLOPs per pair per layer (QKᵀ + PV over all heads).  Fully causal sequences have (T+1)/2 visible keys per token,
    i.e. SURVEY §8's 2dT; a segmented stream has fewer and is credited for fewer."""
    d, L, H = ctor["n_embd"], ctor["n_layer"], ctor["n_head"]
    kvd = d * (ctor.get("n_kv_head") or H) // H
    vis = (T + 1) / 2.0 if visible_keys_per_token is None else float(visible_keys_per_token)
    mlp = 3 * 2 * d * int(8 * d // 3) if ctor.get("use_swiglu") else 16 * d * d
    per_layer = 2 * d * (2 * d + 2 * kvd) + mlp + 4 * d * vis
    n_off = len(ctor.get("multi_offset_targets") or [])
    return L * per_layer + 2 * d * V + n_off * (4 * d * d + 2 * d * V)


def train_flops_per_token(n_layer, d, T, V=68, n_off=5, visible_keys_per_token=None):
    """C3 convenience form: train = 3 x forward."""
    ctor = dict(n_layer=n_layer, n_head=8, n_embd=d, multi_offset_targets=list(range(n_off)))
    return 3 * fwd_flops_per_token(ctor, T, visible_keys_per_token, V)


def visible_keys_per_token(idx, sep_id=3):
    """Mean number of keys a query may attend to under the reference's mask (causal AND same segment, where
    seg = cumsum(idx == sep), model_tiny_gpt.py:273-295): i - start(i) + 1 with start(i) = last <SEP> at or before i."""
    x = idx.numpy()
    B, T = x.shape
    pos = np.broadcast_to(np.arange(T), (B, T))
    start = np.maximum.accumulate(np.where(x == sep_id, pos, 0), axis=1)
    return float((pos - start + 1).mean())


def synthetic_tokens(B, T, seed, vocab=68, kind="random"):
    """kind="random" (the headline, north_star's "synthetic random-codon batches"): codon ids U{4..67},
    no <SEP> — every sequence is one segment, attention is fully causal.  kind="realistic" (second line): BOS first;
    EOS(2)+SEP(3) every U{100..400}; PAD(0) tails on half the rows (exercises segment masks, offset boundary masks,
    termination labels).  targets = ids shifted left, last column PAD."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(4, vocab, size=(B, T), dtype=np.int64)
    for b in range(B if kind == "realistic" else 0):
        idx[b, 0] = 1
        t = int(rng.integers(100, 401))
        while t + 1 < T:
            idx[b, t], idx[b, t + 1] = 2, 3
            t += int(rng.integers(100, 401))
        if rng.random() < 0.5:
            idx[b, T - int(rng.integers(1, T // 2)):] = 0
    tgt = np.zeros_like(idx)
    tgt[:, :-1] = idx[:, 1:]
    return torch.from_numpy(idx), torch.from_numpy(tgt)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val == "Active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def offset_weights(wl):
    return {o: 0.2 for o in wl["offsets"]} if wl["offsets"] else None


def config_dict(args, wl, world, note=None):
    c = {"workload": wl["name"].format(L=args.layers or wl["ctor"]["n_layer"], T=args.seq),
         "per_gpu_batch": args.batch, "seq_len": args.seq, "global_batch": args.batch * world,
         "parallelism": f"dp{world}",
         "tokens": ("random codons U{4..67}, one segment per sequence (full causal attention)" if args.tokens == "random"
                    else "realistic synthetic (BOS, EOS+SEP every U{100..400}, PAD tails on half the rows)"),
         "dropout": args.dropout,
         "l2": "per-step working set (GBs of activations) is far larger than the 126 MB L2; no explicit flush"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference module (vendored under baseline/_ref by build()), or — when it
# is absent — the oracle port of the same PyTorch fp32 path
# ------------------------------------------------------------------------------------------------
def _reference_modules():
    """(TinyGPT, objectives) of the unmodified reference from baseline/_ref, or None."""
    if not os.path.exists(os.path.join(REF_DIR, "src", "codonlm", "model_tiny_gpt.py")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        from src.codonlm.model_tiny_gpt import TinyGPT as RefTinyGPT
        from src.codonlm.training import objectives as ref_obj
    except Exception as exc:  # pragma: no cover
        print(f"[bench] vendored reference not importable ({exc}); using the oracle port", file=sys.stderr)
        return None
    if "codonlm_b200" in getattr(RefTinyGPT, "__module__", ""):
        return None  # the overlay is on the path: that would time this repo, not the reference
    return RefTinyGPT, ref_obj


def _reference_step_fn(wl, ctor, device, autocast_dtype=None):
    """One training step (fwd + the trainer's loss composition loop.py:1067-1143 + bwd + AdamW with the reference's two
    parameter groups loop.py:681-731) of the unmodified reference module -> callable(idx, tgt) -> loss tensor."""
    RefTinyGPT, ref_obj = _reference_modules()
    torch.manual_seed(1337)
    model = RefTinyGPT(**ctor)
    with torch.no_grad():
        model.tok_emb.weight.mul_(0.02)
        if model.pos_emb is not None:
            model.pos_emb.weight.mul_(0.02)
    model = model.to(device).train()
    fast = [p for n, p in model.named_parameters() if ("offset_projs" in n or "termination_head" in n)]
    slow = [p for n, p in model.named_parameters() if not ("offset_projs" in n or "termination_head" in n)]
    groups = ([{"params": fast, "lr": 3e-4, "weight_decay": 0.0}] if fast else []) + \
        [{"params": slow, "lr": 3e-4, "weight_decay": 0.05}]
    opt = torch.optim.AdamW(groups)
    ow = offset_weights(wl)

    def step(idx, tgt):
        with torch.autocast(device_type=device.type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
            if ow:
                logits, loss, aux = model(idx, tgt, return_aux=True)
                off_total, _ = ref_obj.multi_offset_lm_loss(aux["offset_logits"], tgt, ow, label_smoothing=0.05,
                                                            loss_weights=None)
                labels = ref_obj.termination_distance_bucket_labels(tgt, stop_ids=(2,), bucket_edges=(0, 3, 10, 30))
                loss = loss + off_total + TERM_W * ref_obj.termination_aux_loss(aux["termination_logits"], labels)
            else:
                _, loss = model(idx, tgt)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.detach()
    return step


def _port_step_fn(wl, ctor):
    from oracle import codon_gpt_oracle as O  # the one place bench.py touches oracle/ (CPU baseline, reference absent)
    cfg = O.make_cfg(**ctor)
    sd = O.init_state_dict(cfg, seed=1337, emb_scale=0.02)
    leaves = {}
    for k, v in sd.items():
        leaves[k] = v.clone().requires_grad_(True) if (v.dtype.is_floating_point and k != "loss_weights") else v
    leaves["head.weight"] = leaves["tok_emb.weight"]
    params = [v for k, v in leaves.items() if isinstance(v, torch.Tensor) and v.requires_grad and k != "head.weight"]
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=0.05)
    ow = offset_weights(wl)

    def step(idx, tgt):
        total, _, _ = O.training_loss(leaves, cfg, idx, tgt, offset_weights=ow,
                                      termination_loss_weight=TERM_W if ow else 0.0)
        total.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return total.detach()
    return step


def cpu_reference_steps(wl, ctor, seq, steps, warmup, budget_s, batch=2, tokens="random"):
    """The reference's CPU path on all host threads, on a bounded sample (B=`batch` sequences) of the workload."""
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    have_ref = _reference_modules() is not None
    step = _reference_step_fn(wl, ctor, torch.device("cpu")) if have_ref else _port_step_fn(wl, ctor)
    idx, tgt = synthetic_tokens(batch, seq, seed=1337, kind=tokens)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(idx, tgt)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s and time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    toks = batch * seq
    what = ("the unmodified reference module (baseline/_ref: src.codonlm.model_tiny_gpt.TinyGPT + "
            "src.codonlm.training.objectives)" if have_ref else "oracle port of the reference's PyTorch fp32 path")
    return {"value": toks * len(times) / sum(times), "unit": UNIT, "cores": cores,
            "kind": "reference" if have_ref else "port",
            "sample": f"{len(times)} steps of B={batch} x T={seq} ({what}, fp32 torch CPU, AdamW), {warmup} warm-up",
            "ms_per_step": 1e3 * sum(times) / len(times), "steps": len(times), "warmup": warmup}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ctor = model_ctor(wl, args.layers, args.seq, 0.0)
    base = cpu_reference_steps(wl, ctor, args.seq, args.steps, max(0, args.warmup), budget_s=240, tokens=args.tokens)
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": base["steps"],
            "warmup": base["warmup"], "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": config_dict(args, wl, args.gpus, note="reference CPU path on a bounded sample (B=2 sequences per step)"),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def gpu_eager_baseline(wl, ctor, seq, tokens, dev, batch=8, steps=3):
    """The unmodified reference module on THIS B200 in eager PyTorch (fp32 with TF32 off, and bf16 autocast) — the
    number a user of the reference sees on this box before switching (BASELINE.md §4 / SURVEY §8d 'second baseline')."""
    if _reference_modules() is None:
        return {"unavailable": "baseline/_ref is not vendored on this box"}
    out = {"batch": batch, "seq": seq, "steps": steps, "what": "unmodified reference TinyGPT + objectives + torch AdamW, "
           "eager CUDA on the same GPU (SDPA branch with the reference's explicit boolean mask)"}
    idx, tgt = (t.to(dev) for t in synthetic_tokens(batch, seq, seed=1337, kind=tokens))
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for name, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
            step = _reference_step_fn(wl, ctor, dev, dt)
            for _ in range(2):
                step(idx, tgt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                step(idx, tgt)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"tokens_per_s": batch * seq / ms * 1e3, "ms_per_step": ms}
            del step
            torch.cuda.empty_cache()
    except Exception as exc:  # informational line: never fails the bench
        out["error"] = f"{type(exc).__name__}: {exc}"[:300]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return dist, world, rank, local, dev


def _teardown(dist, world, release=None):
    """Release the captured graph before the communicator goes away (a graph that still references NCCL kernels has
    been seen to block communicator destruction).  A teardown that still hangs is reported as a FAILURE (exit 17)."""
    if world <= 1:
        return
    import gc

    def _hung():
        sys.stderr.write("[bench] teardown hung for 60 s\n")
        os._exit(17)
    killer = threading.Timer(60.0, _hung)
    killer.daemon = True
    killer.start()
    if release is not None:
        release()
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    killer.cancel()


def run_train(args, wl):
    dist, world, rank, local, dev = _dist_setup()
    from codonlm_b200 import TinyGPT, ops
    from codonlm_b200.trainer import TrainStep

    ctor = model_ctor(wl, args.layers, args.seq, args.dropout)
    torch.manual_seed(1337)
    model = TinyGPT(**ctor)
    with torch.no_grad():  # trained-scale embeddings: keeps the loss in a realistic range
        model.tok_emb.weight.mul_(0.02)
        if model.pos_emb is not None:
            model.pos_emb.weight.mul_(0.02)
    model = model.to(dev).train()
    ow = offset_weights(wl)
    step = TrainStep(model, lr=3e-4, lr_embedding=3e-4, weight_decay=0.05, offset_weights=ow,
                     termination_loss_weight=TERM_W if ow else 0.0)
    B, T = args.batch, args.seq
    n_host = 4
    host = [synthetic_tokens(B, T, seed=1337 + 1000 * rank + i, kind=args.tokens) for i in range(n_host)]
    vis = float(np.mean([visible_keys_per_token(x) for x, _ in host]))
    pinned = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]

    # ---- GEMM instrumentation (roofline of the dominant kernel), live in the timed region
    gemm_events = []
    orig_gemm = ops.gemm

    def timed_gemm(a, b, out, *, M, N, K, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig_gemm(a, b, out, M=M, N=N, K=K, **kw)
        e.record()
        nbytes = 2.0 * (M * K + N * K) + M * N * out.element_size()
        for extra in ("aux", "aux_out"):
            if kw.get(extra) is not None:
                nbytes += 2.0 * M * N
        if kw.get("residual") is not None or kw.get("accumulate"):
            nbytes += 4.0 * M * N
        gemm_events.append((s, e, 2.0 * M * N * K, nbytes))
        return r

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        loss = step.step(*resident[i % n_host])
    sync_all()
    first_loss = float(loss.item()) if args.warmup else None
    # one eager step counts the kernels this library launches per step (a graph replay launches the same ones)
    l0 = ops.launch_count()
    step.step(*resident[0])
    launches_per_step = ops.launch_count() - l0
    use_graph = not args.no_graph and (world == 1 or os.environ.get("CGPT_BENCH_GRAPH_DDP", "1") == "1")
    if use_graph:
        try:
            step.capture(B, T, allow_collectives=world > 1)
            for i in range(2):
                step.step(*resident[i % n_host])
        except Exception as exc:  # pragma: no cover - capture is an optimisation, never a requirement
            print(f"[bench] CUDA graph capture failed ({exc}); running eagerly", file=sys.stderr)
            step._graph = None
            use_graph = False
    sync_all()

    # ---- timed region 1: device-resident inputs
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    if not use_graph:
        ops.gemm = timed_gemm
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t0.record()
    for i in range(args.steps):
        loss = step.step(*resident[i % n_host])
    t1.record()
    sync_all()
    ops.gemm = orig_gemm
    launches = launches_per_step * args.steps
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    last_loss = float(loss.item())

    # ---- roofline of the dominant kernel family (GEMM): per-launch CUDA events.  With a graph-replayed timed
    # region the events cannot sit inside it, so the same K steps are repeated eagerly with the events on
    # (clock sampler still running); otherwise they were recorded in the timed region itself.
    roof_ms_total = ms_total
    if use_graph:
        graph, step._graph = step._graph, None
        for i in range(2):  # the eager path allocates outside the graph's private pool: let the allocator settle
            step.step(*resident[i % n_host])
        ops.gemm = timed_gemm
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        r0.record()
        for i in range(args.steps):
            step.step(*resident[i % n_host])
        r1.record()
        sync_all()
        ops.gemm = orig_gemm
        roof_ms_total = r0.elapsed_time(r1)
        step._graph = graph
        del graph  # the ONLY owner must be `step`: a captured graph that outlives the communicator blocks its teardown
    clock_info = clocks.stop() if clocks else None

    # ---- timed region 2: end to end through the public call, host buffers in, host loss out
    for i in range(2):
        step.step_host(*pinned[i % n_host])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        step.step_host(*pinned[i % n_host])
    e1.record()
    sync_all()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_ms = float(ems.item())

    if args.breakdown and rank == 0:
        step_breakdown(step, resident[0], args.breakdown, ms_total / args.steps)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        toks_step = B * T * world
        value = toks_step * args.steps / (ms_total / 1e3)
        fpt = 3 * fwd_flops_per_token(ctor, T, visible_keys_per_token=vis)
        fpt_causal = 3 * fwd_flops_per_token(ctor, T)
        gemm_ms = sum(ev[0].elapsed_time(ev[1]) for ev in gemm_events)
        gemm_flops = sum(ev[2] for ev in gemm_events)
        gemm_bytes = sum(ev[3] for ev in gemm_events)
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        traffic = traffic_src = None
        tpath = os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")  # from the committed ncu launch list
        if os.path.exists(tpath) and args.workload == "c3":
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(config_dict(args, wl, world),
                           execution=("whole step (fwd+bwd+AdamW) captured once, replayed as one CUDA graph"
                                      if use_graph else "eager launches from Python")),
            "clocks": clock_info,
            "e2e": {"value": toks_step * args.steps / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": 2 * B * T * 8, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "gemm_bf16_kernel (tcgen05, all instances in the step)", "bound": "tensor",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch (ncu, family average)",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": gemm_bytes / max(1, len(gemm_events)),
                         "peak_source": f"bf16_tflops_sustained, {peak_src}",
                         "launches_per_step": len(gemm_events) // max(1, args.steps),
                         "share_of_step": gemm_ms / roof_ms_total,
                         "timed_in": ("eager instrumented repeat of the K steps right after the timed region "
                                      f"({roof_ms_total / args.steps:.2f} ms/step); the timed region replays a CUDA graph"
                                      if use_graph else "the timed region itself")},
            "step_flops": {"train_flops_per_token": fpt, "model_tflops": value / world * fpt / 1e12,
                           "frac_of_bf16_burst_peak": value / world * fpt / 1e12 / peaks["bf16_tflops"],
                           "frac_of_bf16_sustained_peak": value / world * fpt / 1e12 / peaks["bf16_tflops_sustained"],
                           "attention_visible_keys_per_token": vis, "full_causal_keys_per_token": (T + 1) / 2.0,
                           "note": "attention FLOPs credited for the (query, key) pairs the mask makes visible "
                                   "(executed work); full-causal credit would be %.4g FLOPs/token" % fpt_causal},
            "loss": {"first": first_loss, "last": last_loss},
        }
        if args.workload != "c3":
            line["headline"] = False
        if world == 1 and not args.no_cpu_baseline:
            base = cpu_reference_steps(wl, model_ctor(wl, args.layers, args.seq, 0.0), args.seq, steps=8, warmup=1,
                                       budget_s=25, tokens=args.tokens)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if not args.no_gpu_baseline:
                step._graph = None  # free the graph's private pool for the eager reference
                torch.cuda.empty_cache()
                line["gpu_eager_baseline"] = gpu_eager_baseline(wl, model_ctor(wl, args.layers, args.seq, 0.0), args.seq,
                                                                args.tokens, dev)
        _emit(line)
    _teardown(dist, world, release=lambda: setattr(step, "_graph", None))


def run_infer(args, wl):
    """C4 'batched next-codon inference': last-position logits + argmax for a batch of contexts
    (reference generate.py:14-27 reads logits[:, -1] after a full forward).  value = context tokens/s."""
    dist, world, rank, local, dev = _dist_setup()
    from codonlm_b200 import TinyGPT, ops
    ctor = model_ctor(wl, args.layers, args.seq, 0.0)
    torch.manual_seed(1337)
    model = TinyGPT(**ctor)
    with torch.no_grad():
        model.tok_emb.weight.mul_(0.02)
        model.pos_emb.weight.mul_(0.02)
    model = model.to(dev).eval()
    B, T = args.batch, args.seq
    host = [synthetic_tokens(B, T, seed=11 + 1000 * rank + i, kind=args.tokens)[0] for i in range(4)]
    pinned = [x.pin_memory() for x in host]
    resident = [x.to(dev) for x in host]
    vis = float(np.mean([visible_keys_per_token(x) for x in host]))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            model.next_token_logits(resident[i % 4]).argmax(-1)
        l0 = ops.launch_count()
        model.next_token_logits(resident[0]).argmax(-1)
        launches_per_step = ops.launch_count() - l0
        clocks = ClockSampler(local) if rank == 0 else None
        if clocks:
            clocks.start()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        t0.record()
        for i in range(args.steps):
            model.next_token_logits(resident[i % 4]).argmax(-1)
        t1.record()
        sync_all()
        clock_info = clocks.stop() if clocks else None
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        # end to end: pinned host contexts in, argmax ids back on the host
        buf = torch.empty((B, T), dtype=torch.int64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for i in range(args.steps):
            buf.copy_(pinned[i % 4], non_blocking=True)
            model.next_token_logits(buf).argmax(-1).cpu()
        e1.record()
        sync_all()
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    if rank == 0:
        peaks, peak_src = measured_peaks()
        toks = B * T * world * args.steps
        value = toks / (float(ms.item()) / 1e3)
        f = fwd_flops_per_token(ctor, T, visible_keys_per_token=vis)
        tf = value / world * f / 1e12
        _emit({"metric": "batched next-codon inference, context tokens/sec", "value": value, "unit": UNIT, "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms.item()) / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "headline": False, "config": dict(config_dict(args, wl, world), execution="eager launches from Python"),
               "clocks": clock_info, "contexts_per_s": value / T,
               "e2e": {"value": toks / (float(ems.item()) / 1e3), "unit": UNIT, "h2d_bytes_per_step": B * T * 8,
                       "d2h_bytes_per_step": B * 8},
               "gpu_launches": int(launches_per_step * args.steps),
               "roofline": {"kernel": "whole forward (tcgen05 GEMMs + attention)", "bound": "tensor", "achieved": tf,
                            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
                            "traffic": None, "peak_source": f"bf16_tflops (burst), {peak_src}"},
               "step_flops": {"fwd_flops_per_token": f, "model_tflops": tf}})
    _teardown(dist, world)


def run_attn(args, wl):
    """C5: causal attention forward + backward only (H=8, hd=64), seq 4096, B sequences per GPU; the ranks are
    independent replicas over sequences (no collective inside attention).  value = tokens/s through one
    attention layer fwd+bwd."""
    dist, world, rank, local, dev = _dist_setup()
    from codonlm_b200 import ops
    B, T, H, hd = args.batch, args.seq, 8, 64
    g = torch.Generator(device="cpu").manual_seed(5 + rank)
    qkv = torch.randn(B * T, 3 * H * hd, generator=g).to(torch.bfloat16).to(dev)
    ss = None
    if args.tokens == "realistic":
        ss = ops.segment_starts(synthetic_tokens(B, T, seed=5 + rank, kind="realistic")[0].to(dev), 3)
    out, lse = ops.attn_fwd(qkv, ss, B, T, H, H, hd)
    dout = torch.randn_like(out)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one():
        o, l = ops.attn_fwd(qkv, ss, B, T, H, H, hd)
        return ops.attn_bwd(qkv, ss, o, dout, l, B, T, H, H, hd)

    for _ in range(args.warmup):
        one()
    l0 = ops.launch_count()
    one()
    launches_per_step = ops.launch_count() - l0
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t0.record()
    for _ in range(args.steps):
        one()
    t1.record()
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        ops.attn_fwd(qkv, ss, B, T, H, H, hd)
    f1.record()
    sync_all()
    clock_info = clocks.stop() if clocks else None
    ms = torch.tensor([t0.elapsed_time(t1), f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        peaks, peak_src = measured_peaks()
        if ss is None:
            pairs = B * H * T * (T + 1) // 2
        else:
            pos = torch.arange(T, device=dev)[None, :]
            pairs = int((pos - ss.long() + 1).sum().item()) * H
        ms_fb, ms_f = float(ms[0].item()) / args.steps, float(ms[1].item()) / args.steps
        tf = 14 * hd * pairs / ms_fb / 1e9  # 4·hd forward + 10·hd backward FLOPs per visible pair
        _emit({"metric": "causal attention fwd+bwd tokens/sec (one layer)", "value": B * T * world / ms_fb * 1e3, "unit": UNIT,
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_fb, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "headline": False,
               "config": dict(config_dict(args, wl, world), execution="eager launches from Python; replicas over sequences"),
               "clocks": clock_info, "gpu_launches": int(launches_per_step * args.steps),
               "us_forward": ms_f * 1e3, "us_backward": (ms_fb - ms_f) * 1e3,
               "e2e": None,
               "roofline": {"kernel": "attn_fwd_w3_kernel<64> + attn_bwd_ws_kernel<64> (+ delta / dQ convert)",
                            "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": tf / peaks["bf16_tflops"], "traffic": None, "visible_pairs": pairs,
                            "tflops_forward": 4 * hd * pairs / ms_f / 1e9,
                            "tflops_backward": 10 * hd * pairs / max(ms_fb - ms_f, 1e-9) / 1e9,
                            "peak_source": f"bf16_tflops (burst), {peak_src}"}})
    _teardown(dist, world)


def step_breakdown(step, batch, path, ms_step):
    """Extra (untimed) steps with every C-ABI wrapper bracketed by CUDA events: per-op, per-shape device time.
    Diagnostic only: the numbers bench.py reports never come from here."""
    from codonlm_b200 import ops
    names = ["segment_starts", "next_in_set", "termination_labels", "embed_fwd", "embed_bwd", "layernorm_fwd",
             "layernorm_bwd", "gemm", "cast_bf16", "colsum_bf16", "rope_qk", "swiglu_fwd", "swiglu_bwd", "attn_fwd",
             "attn_bwd", "skinny_linear_fwd", "skinny_linear_bwd", "ce_fwd", "ce_bwd", "adamw", "split3", "dropout"]
    rec, orig = [], {}

    def wrap(name, fn):
        def w(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            if name == "gemm":
                sig = f"M{k['M']} N{k['N']} K{k['K']} a_mn{int(k.get('a_mn', False))} b_mn{int(k.get('b_mn', False))} " \
                      f"split{k.get('split_k', 1)} epi{k.get('epilogue', 0)} res{int(k.get('residual') is not None)} " \
                      f"{'f32' if a[2].dtype == torch.float32 else 'bf16'}"
                fl = 2.0 * k["M"] * k["N"] * k["K"]
            else:
                t0 = next((x for x in a if isinstance(x, torch.Tensor)), None)
                sig = "x".join(str(d) for d in t0.shape) if t0 is not None else ""
                fl = 0.0
            rec.append((name, sig, s, e, fl))
            return r
        return w
    for n in names:
        if hasattr(ops, n):
            orig[n] = getattr(ops, n)
            setattr(ops, n, wrap(n, orig[n]))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nsteps = 2
    torch.cuda.synchronize()
    t0.record()
    for _ in range(nsteps):
        step._eager_step(*batch)  # eager: a CUDA-graph replay would bypass the wrapped ops
    t1.record()
    torch.cuda.synchronize()
    for n in orig:
        setattr(ops, n, orig[n])
    agg = {}
    for name, sig, s, e, fl in rec:
        a = agg.setdefault((name, sig), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += s.elapsed_time(e)
        a[2] += fl
    total = t0.elapsed_time(t1) / nsteps
    rows = sorted(((k, v) for k, v in agg.items()), key=lambda kv: -kv[1][1])
    covered = sum(v[1] for _, v in rows) / nsteps
    with open(path, "w") as f:
        f.write(f"# one training step: {total:.2f} ms with per-op events ({ms_step:.2f} ms in the timed region); "
                f"ops below cover {covered:.2f} ms\n# op | shape | launches/step | ms/step | share | TFLOP/s\n")
        for (name, sig), (cnt, ms, fl) in rows:
            tf = f"{fl / (ms / 1e3) / 1e12:8.1f}" if fl > 0 and ms > 0 else "       -"
            f.write(f"{name:18s} | {sig:70s} | {cnt / nsteps:6.1f} | {ms / nsteps:8.3f} | {ms / nsteps / total:6.1%} | {tf}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS),
                    help="c3 = the headline (BASELINE.json configs[2]); the others are the remaining configs")
    ap.add_argument("--batch", type=int, default=None, help="sequences per GPU per step (default: the workload's)")
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--tokens", default="random", choices=["random", "realistic"],
                    help="random = north_star's random-codon batches (headline); realistic = segmented stream with PAD tails")
    ap.add_argument("--dropout", type=float, default=0.0, help="dropout probability of the model (reference configs: 0.1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-reference-on-this-GPU baseline")
    ap.add_argument("--no-graph", action="store_true", help="run the timed region eagerly (default: CUDA graph)")
    ap.add_argument("--breakdown", default=None, help="write a per-op CUDA-event breakdown of one step to this file")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.batch = args.batch or wl["batch"]
    args.seq = args.seq or wl["seq"]
    if args.impl == "reference":
        if wl["kind"] != "train":
            wl = WORKLOADS["c3"]
        run_reference(args, wl)
        return
    if args.warmup < 3:
        args.warmup = 3
    {"train": run_train, "infer": run_infer, "attn": run_attn}[wl["kind"]](args, wl)


def _emit(line: dict):
    """The ONE JSON line goes to the process's real stdout; everything else any library prints while the bench
    runs (NCCL's version banner, torchrun warnings, ...) was redirected to stderr by _guard_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _guard_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run (C libraries included)


if __name__ == "__main__":
    _guard_stdout()
    main()
