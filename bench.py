#!/usr/bin/env python
"""Benchmark of the codon-GPT training step (BASELINE.json metric: training codon tokens/sec).

  python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W # reference CPU path (oracle port), rank 0
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # data-parallel

Workload (config.workload): BASELINE.json configs[2] = "12L8H d512 separate/multi-offset heads training
step, seq 1024" (the configuration the metric is quoted on; it fits one GPU): per-GPU micro-batch of 64
sequences x 1024 codons, offsets [2,4,8,16,32] + termination head, label smoothing 0.05, AdamW.
A step = forward + backward (+ gradient all-reduce) + optimiser update on one synthetic batch.
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

METRIC = "training codon tokens/sec"
UNIT = "tokens/s"
OFFSETS = [2, 4, 8, 16, 32]
OFFSET_W = {o: 0.2 for o in OFFSETS}
TERM_W = 0.1


def workload_ctor(n_layer=12, seq=1024, dropout=0.0):
    return dict(vocab_size=68, block_size=seq, n_layer=n_layer, n_head=8, n_embd=512, dropout=dropout,
                label_smoothing=0.05, sep_id=3, use_sdpa=True, termination_aux=True, multi_offset_targets=OFFSETS)


def train_flops_per_token(n_layer, d, T, V=68, n_off=5, visible_keys_per_token=None):
    """Algorithmic FLOPs per token, train = 3 x fwd.  Attention is credited for the (query, key) pairs the mask makes
    visible: 4·d FLOPs per pair per layer forward (QKᵀ + PV over all heads).  Fully causal sequences have (T+1)/2
    visible keys per token, i.e. SURVEY §8's 2dT; a segmented stream has fewer and is credited for fewer."""
    vis = (T + 1) / 2.0 if visible_keys_per_token is None else float(visible_keys_per_token)
    fwd = n_layer * (24 * d * d + 4 * d * vis) + 2 * d * V + n_off * (4 * d * d + 2 * d * V)
    return 3 * fwd


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


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's PyTorch fp32 CPU path
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(n_layer, seq, steps, warmup, budget_s, batch=2, tokens="random"):
    from oracle import codon_gpt_oracle as O  # the one place bench.py touches oracle/ (CPU baseline legs)
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    ctor = workload_ctor(n_layer, seq)
    cfg = O.make_cfg(**ctor)
    sd = O.init_state_dict(cfg, seed=1337, emb_scale=0.02)
    leaves = {}
    for k, v in sd.items():
        leaves[k] = v.clone().requires_grad_(True) if (v.dtype.is_floating_point and k != "loss_weights") else v
    leaves["head.weight"] = leaves["tok_emb.weight"]
    params = [v for k, v in leaves.items() if isinstance(v, torch.Tensor) and v.requires_grad and k != "head.weight"]
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=0.05)
    idx, tgt = synthetic_tokens(batch, seq, seed=1337, kind=tokens)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        total, _, _ = O.training_loss(leaves, cfg, idx, tgt, offset_weights=OFFSET_W, termination_loss_weight=TERM_W)
        total.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s and time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    toks = batch * seq
    return {"value": toks * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} steps of B={batch} x T={seq} (same model, fp32 torch CPU, AdamW), "
                      f"{warmup} warm-up", "ms_per_step": 1e3 * sum(times) / len(times), "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference_steps(args.layers, args.seq, args.steps, max(1, min(args.warmup, 2)), budget_s=240,
                               tokens=args.tokens)
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": base["steps"],
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": config_dict(args, args.gpus, note="reference CPU path on a bounded sample (B=2)"),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def config_dict(args, world, note=None):
    c = {"workload": f"C3 codon-GPT {args.layers}L8H d512 (MHA hd64, GELU 2048) + offset heads {OFFSETS} + "
                     f"termination head, train step fwd+bwd+AdamW, seq {args.seq}",
         "per_gpu_batch": args.batch, "seq_len": args.seq, "global_batch": args.batch * world,
         "parallelism": f"dp{world}",
         "tokens": ("random codons U{4..67}, one segment per sequence (full causal attention)" if args.tokens == "random"
                    else "realistic synthetic (BOS, EOS+SEP every U{100..400}, PAD tails on half the rows)"),
         "dropout": args.dropout,
         "l2": "per-step working set (~14 GB of activations) is far larger than the 126 MB L2; no explicit flush"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
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
    from codonlm_b200 import TinyGPT, ops
    from codonlm_b200.trainer import TrainStep

    torch.manual_seed(1337)
    model = TinyGPT(**workload_ctor(args.layers, args.seq, args.dropout))
    with torch.no_grad():  # trained-scale embeddings: keeps the loss in a realistic range
        model.tok_emb.weight.mul_(0.02)
        model.pos_emb.weight.mul_(0.02)
    model = model.to(dev).train()
    step = TrainStep(model, lr=3e-4, lr_embedding=3e-4, weight_decay=0.05, offset_weights=OFFSET_W,
                     termination_loss_weight=TERM_W)
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
        fpt = train_flops_per_token(args.layers, 512, T, visible_keys_per_token=vis)
        fpt_causal = train_flops_per_token(args.layers, 512, T)
        gemm_ms = sum(ev[0].elapsed_time(ev[1]) for ev in gemm_events)
        gemm_flops = sum(ev[2] for ev in gemm_events)
        gemm_bytes = sum(ev[3] for ev in gemm_events)
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        traffic = traffic_src = None
        tpath = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")  # from the committed ncu launch list
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(config_dict(args, world),
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
        if world == 1 and not args.no_cpu_baseline:
            base = cpu_reference_steps(args.layers, args.seq, steps=8, warmup=1, budget_s=25, tokens=args.tokens)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        _emit(line)
    if world > 1:
        # tear down with a watchdog: destroying the NCCL communicator while a captured graph still references its
        # kernels has been seen to block; the JSON line is out, so a stuck teardown must not keep the job alive
        killer = threading.Timer(20.0, os._exit, args=(0,))
        killer.daemon = True
        killer.start()
        step._graph = None
        torch.cuda.synchronize()
        dist.destroy_process_group()
        killer.cancel()


def step_breakdown(step, batch, path, ms_step):
    """Extra (untimed) steps with every C-ABI wrapper bracketed by CUDA events: per-op, per-shape device time.
    Diagnostic only: the numbers bench.py reports never come from here."""
    from codonlm_b200 import ops
    names = ["segment_starts", "next_in_set", "termination_labels", "embed_fwd", "embed_bwd", "layernorm_fwd",
             "layernorm_bwd", "gemm", "cast_bf16", "colsum_bf16", "rope_qk", "swiglu_fwd", "swiglu_bwd", "attn_fwd",
             "attn_bwd", "skinny_linear_fwd", "skinny_linear_bwd", "ce_fwd", "ce_bwd", "adamw"]
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
    for n in names:
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
    ap.add_argument("--batch", type=int, default=64, help="sequences per GPU per step")
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--tokens", default="random", choices=["random", "realistic"],
                    help="random = north_star's random-codon batches (headline); realistic = segmented stream with PAD tails")
    ap.add_argument("--dropout", type=float, default=0.0, help="dropout probability of the model (reference configs: 0.1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the timed region eagerly (default: CUDA graph at N=1)")
    ap.add_argument("--breakdown", default=None, help="write a per-op CUDA-event breakdown of one step to this file")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


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
