"""The training step around the model: flat fp32 parameter / gradient buffers, fused AdamW, and the
data-parallel gradient exchange (bucketed bf16 all-reduce over NCCL, overlapped with backward).

Host-side mirror of what the reference trainer does per optimiser step
(src/codonlm/training/loop.py: fwd() :1067-1143, step_optimizer :1145-1182,
_average_accumulated_gradients :145-150, AdamW param groups :681-731) minus its four host syncs per
micro-batch: losses stay on the device and are read back only when the caller asks.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import functional as Fn
from . import ops
from .model_tiny_gpt import bump_shadow_generation
from .objectives import training_loss

HEAD_GROUP_MARKERS = ("shape_proj", "offset_projs", "termination_head")  # loop.py:660-667,689


def split_param_groups(model) -> Dict[str, List[tuple]]:
    """The reference's two AdamW groups (loop.py:681-731): the aux heads get lr_embedding and no weight decay,
    everything else (tok_emb, LayerNorms and biases included — the "transformer.wte" test at :689 never
    matches) is decayed.  Order inside a group = reverse execution order, so that gradient buckets fill from
    the front while backward runs."""
    seen = set()
    named = []
    for name, p in model.named_parameters():
        if id(p) in seen or not p.requires_grad:
            continue
        seen.add(id(p))
        named.append((name, p))
    named.reverse()
    named = _packed_order(model, named)
    groups = {"head": [], "backbone": []}
    for name, p in named:
        groups["head" if any(m in name for m in HEAD_GROUP_MARKERS) else "backbone"].append((name, p))
    return groups


def _packed_order(model, named):
    """Parameters that a module evaluates as ONE packed operand (query|key|value weights and biases,
    CausalSelfAttention.packed_param_groups) are made neighbours, in packed order, so that their slices of the
    flat master / shadow / gradient buffers form one contiguous matrix: no per-step packing copies, one wgrad."""
    packs = []
    for m in model.modules():
        f = getattr(m, "packed_param_groups", None)
        if f is not None:
            packs.extend(tuple(g) for g in f())
    name_of = {id(p): n for n, p in named}
    member = {}
    for gi, g in enumerate(packs):
        if all(id(p) in name_of for p in g):
            for p in g:
                member[id(p)] = gi
    out, emitted = [], set()
    for n, p in named:
        gi = member.get(id(p))
        if gi is None:
            out.append((n, p))
        elif gi not in emitted:
            emitted.add(gi)
            out.extend((name_of[id(q)], q) for q in packs[gi])
    return out


class FlatGroup:
    """Parameters of one optimiser group re-homed into a flat fp32 buffer, with flat grad / m / v."""

    def __init__(self, named_params: Sequence[tuple], lr: float, weight_decay: float):
        self.names = [n for n, _ in named_params]
        self.params = [p for _, p in named_params]
        self.lr, self.weight_decay = lr, weight_decay
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.offsets, off = [], 0
        for s in sizes:
            self.offsets.append(off)
            off += (s + 63) // 64 * 64  # keep every parameter 256-byte aligned
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.v = torch.zeros(off, dtype=torch.float32, device=dev)
        # bf16 copy of the masters, the tensor-core operand: written by the AdamW kernel together with the master,
        # so a training step contains no cast kernels at all (model_tiny_gpt.flat_shadow hands out views of it)
        self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = None
                # kernels accumulate straight into this slot (functional._main_grad); autograd never sees a grad
                p.main_grad = self.grad[o:o + p.numel()].view_as(p)
                p._cgpt_shadow = self.shadow[o:o + p.numel()].view_as(p)
                p._cgpt_shadow_version = p._version
            self.shadow.copy_(self.flat)


_PROBE_NO_ALLREDUCE = os.environ.get("CGPT_PROBE_NO_ALLREDUCE", "0") == "1"


class _NoWork:
    """Stand-in for a collective's work handle (CGPT_PROBE_NO_ALLREDUCE=1): wait() orders the streams, nothing more."""

    def __init__(self, stream):
        self.ev = torch.cuda.Event()
        self.ev.record(stream)

    def wait(self):
        torch.cuda.current_stream().wait_event(self.ev)


class GradBuckets:
    """Bucketed bf16 all-reduce of a flat gradient buffer, launched on a side stream as soon as every
    parameter of a bucket has received its gradient."""

    def __init__(self, group: FlatGroup, process_group, bucket_bytes: int = 25 << 20, overlap: bool = True,
                 copy_back: bool = True):
        self.g = group
        # copy_back=False: finish() leaves the reduced gradient in the bf16 staging buffer (`reduced`) and the fused
        # AdamW reads it from there (cgpt_adamw_bf16grad): one pass over the gradients less per step
        self.copy_back = bool(copy_back)
        # overlap=False: every bucket is reduced after backward instead of from the gradient-ready hooks.  The NCCL
        # kernels then never share the SMs with the persistent GEMM / attention kernels (whose grids are sized to all
        # 148 SMs: a CTA that cannot be placed next to an NCCL CTA runs as a second wave)
        self.overlap = bool(overlap)
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.on_gpu = group.grad.is_cuda  # CPU tensors only in the gloo tests of this host logic
        self.comm_stream = torch.cuda.Stream() if self.on_gpu else None
        elems = max(1, bucket_bytes // 2)
        self.bounds = []     # (start, end) element ranges of the flat buffer, in backward-completion order
        self.bucket_of = {}  # param index -> bucket index
        self.need = []       # params per bucket
        start, count = 0, 0
        for i, (p, o) in enumerate(zip(group.params, group.offsets)):
            end = o + (p.numel() + 63) // 64 * 64
            self.bucket_of[i] = len(self.bounds)
            count += 1
            if end - start >= elems or i == len(group.params) - 1:
                self.bounds.append((start, end))
                self.need.append(count)
                start, count = end, 0
        self.left = list(self.need)
        self.reduced = torch.zeros(group.numel, dtype=torch.bfloat16, device=group.grad.device)  # all buckets, flat
        self.staging = [self.reduced[s:e] for s, e in self.bounds]
        self.pending = []
        # A parameter can receive several contributions per backward (tied head/tok_emb, offset heads sharing
        # the head).  The first step counts them (no overlap: everything is reduced in finish()); later steps
        # launch a bucket as soon as all contributions of all its parameters have arrived.
        self.contrib_seen = [0] * len(group.params)
        self.contrib_need = None
        self.last_order = None  # bucket launch order of the last complete step
        self.last_hook_order = None  # ... and the part of it that the gradient-ready hooks launched during backward
        self.armed = True  # False while accumulating the non-final micro-batches of a group: nothing is launched
        for i, p in enumerate(group.params):
            p._cgpt_grad_ready = self._make_hook(i)
            p.register_post_accumulate_grad_hook(lambda _p, i=i: self._make_hook(i)())  # plain-autograd path

    def _make_hook(self, i):
        def hook():
            if not self.armed:
                return
            self.contrib_seen[i] += 1
            if self.contrib_need is None:
                return
            if self.overlap and self.contrib_seen[i] == self.contrib_need[i]:
                b = self.bucket_of[i]
                self.left[b] -= 1
                if self.left[b] == 0:
                    self._launch(b)
        return hook

    def _launch(self, b):
        s, e = self.bounds[b]
        if not self.on_gpu:
            self.staging[b].copy_(self.g.grad[s:e])
            work = dist.all_reduce(self.staging[b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self.pending.append((b, work))
            return
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            stg = self.staging[b]
            ops.cast_bf16(self.g.grad[s:e].view(1, -1), out=stg.view(1, -1))
            if _PROBE_NO_ALLREDUCE:  # measurement only: what the step costs on N ranks without the collective
                work = _NoWork(self.comm_stream)
            else:
                work = dist.all_reduce(stg, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self.pending.append((b, work))

    def join_without_backward(self):
        """A rank that holds no micro-batch of an accumulation group (ragged tail) still has to issue the group's
        collectives, and in the order its peers do: they launch buckets from their gradient-ready hooks DURING backward,
        i.e. before anything the caller all-reduces after backward (the non-finite flags).  Launch the same buckets, in
        the same order, now (the gradients are zero here); finish() adds the rest."""
        if self.last_hook_order and not self.pending:
            for b in self.last_hook_order:
                self._launch(b)

    def finish(self):
        """Wait for all buckets and write the reduced gradients (sum over ranks) back as fp32."""
        launched = {b for b, _ in self.pending}
        if self.contrib_need is not None and self.armed:
            self.last_hook_order = [b for b, _ in self.pending]  # what the gradient-ready hooks launched, in order
        # Buckets not launched from the hooks (first step; parameters without a gradient this step; a rank that held
        # no micro-batch of a ragged accumulation group) are launched here.  Collectives must be issued in the SAME
        # order on every rank: the order the hooks produced in the last complete step, when there is one.
        order = self.last_order if self.last_order is not None else range(len(self.bounds))
        for b in order:
            if b not in launched:
                self._launch(b)
        if len(self.pending) == len(self.bounds):
            self.last_order = [b for b, _ in self.pending]
        if self.contrib_need is None:
            self.contrib_need = [max(1, c) for c in self.contrib_seen]
        self.contrib_seen = [0] * len(self.contrib_seen)
        for b, work in self.pending:
            work.wait()  # the compute stream now waits for that bucket's all-reduce
            if self.copy_back:
                s, e = self.bounds[b]
                self.g.grad[s:e].copy_(self.staging[b])
        self.pending.clear()
        self.left = list(self.need)


class TrainStep:
    """One optimiser step = forward + backward (+ all-reduce) + AdamW on a (B,T) batch of codon ids."""

    HYPER_RING = 8  # pinned staging rows for the per-step AdamW scalars

    def __init__(self, model, lr=3e-4, lr_embedding=None, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8,
                 offset_weights: Optional[Dict[int, float]] = None, termination_loss_weight: float = 0.0,
                 process_group=None, bucket_mb: int = 25, overlap_allreduce: Optional[bool] = None,
                 replay_loss_weight: float = 0.1, replay_class_weights=None, termination_stop_ids=(2,),
                 termination_bucket_edges=(0, 3, 10, 30), termination_class_weights=None):
        self.model = model
        self.termination_kwargs = dict(termination_stop_ids=tuple(termination_stop_ids),
                                       termination_bucket_edges=tuple(termination_bucket_edges),
                                       termination_class_weights=termination_class_weights)
        self.offset_weights = offset_weights
        self.termination_loss_weight = termination_loss_weight
        self.replay_loss_weight = replay_loss_weight
        self.replay_class_weights = replay_class_weights
        self.betas, self.eps = betas, eps
        groups = split_param_groups(model)
        self.groups: List[FlatGroup] = []
        # a group exists only when it has parameters (loop.py:712-726 `if embedding_params:` / `if backbone_params:`):
        # with freeze_backbone (loop.py:656-667) only the aux heads train and the backbone group is empty
        if groups["head"]:
            self.groups.append(FlatGroup(groups["head"], lr_embedding if lr_embedding is not None else lr, 0.0))
        if groups["backbone"]:
            self.groups.append(FlatGroup(groups["backbone"], lr, weight_decay))
        if not self.groups:
            raise ValueError("TrainStep: the model has no trainable parameters")
        self.group_kinds = [k for k in ("head", "backbone") if groups[k]]
        self.step_count = 0
        self.world = 1
        self.buckets: List[GradBuckets] = []
        if process_group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            pg = process_group if process_group is not None else dist.group.WORLD
            self.world = dist.get_world_size(pg)
            if overlap_allreduce is None:
                overlap_allreduce = os.environ.get("CGPT_DDP_OVERLAP", "1") == "1"
            on_gpu = self.groups[0].grad.is_cuda
            self.buckets = [GradBuckets(g, pg, bucket_mb << 20, overlap=overlap_allreduce, copy_back=not on_gpu)
                            for g in self.groups]
        dev = self.groups[0].flat.device
        self._xb = self._yb = None
        self._dev = dev
        # step-dependent AdamW scalars on the device, one row per group: [lr, 1-b1^t, sqrt(1-b2^t)] (graph replay)
        self._hyper = torch.zeros((len(self.groups), 3), dtype=torch.float32, device=dev)
        # The host runs ahead of the stream (nothing in step() synchronises), so ONE pinned staging row would be
        # overwritten for step N+k before the async copy of step N has read it.  A ring of pinned rows, each guarded by
        # an event recorded right after its copy: a slot is rewritten only once its copy has executed.
        self._hyper_ring = [(torch.zeros((len(self.groups), 3), dtype=torch.float32).pin_memory() if dev.type == "cuda"
                             else torch.zeros((len(self.groups), 3)), None) for _ in range(self.HYPER_RING)]
        self._hyper_slot = 0
        self._graph = None
        self.last_lr_scale = 1.0
        # dropout masks: generator state on the device, so that a captured step draws fresh masks on every replay
        self._gen = None
        if dev.type == "cuda" and getattr(model, "dropout_p", 0.0) > 0.0:
            self._gen = ops.DeviceGenerator(dev)
            if dist.is_available() and dist.is_initialized():  # different masks on different ranks
                self._gen.state[1] += dist.get_rank() << 40

    # ------------------------------------------------------------------ pieces
    def zero_grad(self):
        for g in self.groups:
            g.grad.zero_()
            Fn.clear_credits(g.params)  # leftovers of an interrupted backward (functional._credit)

    def forward_backward(self, xb, yb, replay=None):
        with ops.device_generator(self._gen):  # no-op scope without dropout
            total, parts, _ = training_loss(self.model, xb, yb, offset_weights=self.offset_weights,
                                            termination_loss_weight=self.termination_loss_weight, replay=replay,
                                            replay_loss_weight=self.replay_loss_weight,
                                            replay_class_weights=self.replay_class_weights, **self.termination_kwargs)
            total.backward()
            if self._gen is not None:
                self._gen.advance()  # device-side: the next micro-batch (or graph replay) draws new masks
        return total.detach(), parts

    def arm_collectives(self, armed: bool):
        """Gradient buckets are all-reduced only from the backward of a group's LAST micro-batch."""
        for bk in self.buckets:
            bk.armed = bool(armed)

    def join_collectives_without_backward(self):
        for bk in self.buckets:
            bk.join_without_backward()

    def discard_gradients(self):
        """Drop an aborted accumulation group: pending bucket reductions are drained, gradients zeroed."""
        for bk in self.buckets:
            bk.finish()
        self.zero_grad()

    def optimizer_step(self, lr_scale: float = 1.0, micro_batches: int = 1, global_micro_batches: Optional[int] = None):
        """AdamW on the summed gradients divided by the number of micro-batches that produced them (loop.py:145-150):
        `micro_batches` local ones on each of `world` ranks, or — when the ranks hold different numbers (a ragged
        trailing group) — `global_micro_batches`, the sum over the ranks, which every rank must pass identically."""
        self.step_count += 1
        n_global = int(global_micro_batches) if global_micro_batches is not None else micro_batches * self.world
        gscale = 1.0 / max(1, n_global)
        for bk in self.buckets:
            bk.finish()
        self._push_hyper(lr_scale)
        for gi, g in enumerate(self.groups):
            ops.adamw(g.flat, self._grad_source(gi), g.m, g.v, g.shadow if g.shadow.is_cuda else None, g.lr * lr_scale,
                      self.betas[0], self.betas[1], self.eps, g.weight_decay, self.step_count, gscale,
                      dev_hyper=self._hyper[gi])
        # the kernel wrote the masters behind autograd's back: invalidate the bf16 shadow caches
        bump_shadow_generation()

    def _grad_source(self, gi: int):
        """What AdamW reads: the flat fp32 gradient, or — data-parallel on the GPU — the all-reduced bf16 buckets."""
        if self.buckets and not self.buckets[gi].copy_back:
            return self.buckets[gi].reduced
        return self.groups[gi].grad

    def _push_hyper(self, lr_scale: float):
        t = self.step_count
        slot = self._hyper_slot
        self._hyper_slot = (slot + 1) % self.HYPER_RING
        row, ev = self._hyper_ring[slot]
        if ev is not None:
            ev.synchronize()  # the copy that last read this row has executed (HYPER_RING steps ago: never waits in practice)
        for gi, g in enumerate(self.groups):
            row[gi, 0] = g.lr * lr_scale
            row[gi, 1] = 1.0 - self.betas[0] ** t
            row[gi, 2] = math.sqrt(1.0 - self.betas[1] ** t)
        self._hyper.copy_(row, non_blocking=True)
        if self._hyper.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._hyper_ring[slot] = (row, ev)
        self.last_lr_scale = float(lr_scale)

    # ------------------------------------------------------------------ optimiser state in torch.optim.AdamW layout
    def reference_param_groups(self):
        """The reference's optimiser param groups (loop.py:681-726) over this model's parameters: 'embedding' group
        (aux heads: lr_embedding, no decay) first when non-empty, then the backbone, each in named_parameters() order
        — the order that fixes the integer parameter ids of torch's optimizer.state_dict()."""
        fast, slow, seen = [], [], set()
        for name, p in self.model.named_parameters():
            if id(p) in seen or not p.requires_grad:
                continue
            seen.add(id(p))
            (fast if any(m in name for m in HEAD_GROUP_MARKERS) else slow).append(p)
        out = []
        by_kind = dict(zip(self.group_kinds, self.groups))
        if fast:
            g = by_kind["head"]
            out.append({"params": fast, "lr": g.lr, "weight_decay": g.weight_decay})
        if slow:
            g = by_kind["backbone"]
            out.append({"params": slow, "lr": g.lr, "weight_decay": g.weight_decay})
        return out

    def _slot(self, p):
        for g in self.groups:
            for q, o in zip(g.params, g.offsets):
                if q is p:
                    return g, o
        raise KeyError("parameter is not owned by this TrainStep")

    def _shell_optimizer(self):
        """A torch.optim.AdamW over the same parameters and groups, used ONLY as the (de)serialiser of its own
        state_dict layout (it never steps): whatever keys this torch version writes / validates are produced / checked
        by torch itself."""
        opt = torch.optim.AdamW(self.reference_param_groups(), betas=self.betas, eps=self.eps)
        for pg in opt.param_groups:
            pg.setdefault("initial_lr", pg["lr"])  # what LambdaLR adds on construction (loop.py:780)
        return opt

    def state_dict(self, lr_scale_fn=None) -> dict:
        """Optimiser state as `torch.optim.AdamW.state_dict()` of the reference's optimiser would hold it
        (`payload["optimizer"]`, loop.py:962): per-parameter `step`, `exp_avg`, `exp_avg_sq` sliced out of the flat
        m / v buffers, param groups with the current (scheduled) lr.  A reference `last.pt` therefore resumes here and
        a checkpoint written here resumes in the reference (loop.py:891-893).  With `lr_scale_fn` (the LambdaLR lambda)
        the stored lr is the one the reference's scheduler has already set for the NEXT step (it steps right after the
        optimiser, loop.py:1181-1182); without it, the scale of the last step taken."""
        opt = self._shell_optimizer()
        if self.step_count > 0:
            for pg in opt.param_groups:
                for p in pg["params"]:
                    g, o = self._slot(p)
                    n = p.numel()
                    opt.state[p] = {"step": torch.tensor(float(self.step_count)),
                                    "exp_avg": g.m[o:o + n].view_as(p).clone(),
                                    "exp_avg_sq": g.v[o:o + n].view_as(p).clone()}
        scale = float(lr_scale_fn(self.step_count)) if lr_scale_fn is not None else self.last_lr_scale
        for pg in opt.param_groups:
            pg["lr"] = pg["initial_lr"] * scale
        return opt.state_dict()

    def load_state_dict(self, state: dict):
        """Accepts `torch.optim.AdamW.state_dict()` of the reference's optimiser (same two groups, same order)."""
        opt = self._shell_optimizer()
        opt.load_state_dict(state)  # torch validates group count and sizes and casts / moves the state tensors
        steps = set()
        with torch.no_grad():
            for pg in opt.param_groups:
                for p in pg["params"]:
                    st = opt.state.get(p)
                    g, o = self._slot(p)
                    n = p.numel()
                    if not st:
                        g.m[o:o + n].zero_()
                        g.v[o:o + n].zero_()
                        continue
                    g.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    g.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(round(float(st["step"]))))
        if len(steps) > 1:
            raise ValueError(f"optimizer state holds different step counts per parameter {sorted(steps)}: the fused "
                             "AdamW keeps one step counter (the reference steps all parameters together)")
        self.step_count = steps.pop() if steps else 0
        by_kind = dict(zip(self.group_kinds, self.groups))
        kinds = [k for k in ("head", "backbone") if k in by_kind]
        for kind, pg in zip(kinds, opt.param_groups):
            g = by_kind[kind]
            g.weight_decay = float(pg["weight_decay"])
            base = float(pg.get("initial_lr", pg["lr"]))
            g.lr = base
            self.last_lr_scale = float(pg["lr"]) / base if base > 0 else 1.0
        if opt.param_groups:
            self.betas = tuple(opt.param_groups[0]["betas"])
            self.eps = float(opt.param_groups[0]["eps"])
        if self._graph is not None:  # group scalars are baked into a captured step
            self._graph = None

    def scheduler_state_dict(self, lr_scale_fn=None) -> dict:
        """`payload["scheduler"]` of the reference (LambdaLR.state_dict(), loop.py:963) after `step_count` optimiser
        steps: produced by a real LambdaLR over the shell optimiser so the key set is torch's own."""
        opt = self._shell_optimizer()
        fn = lr_scale_fn if lr_scale_fn is not None else (lambda i: 1.0)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, fn)
        sd = sched.state_dict()
        sd["last_epoch"] = int(self.step_count)
        sd["_step_count"] = int(self.step_count) + 1
        sd["_last_lr"] = [pg["initial_lr"] * float(fn(self.step_count)) for pg in opt.param_groups]
        return sd

    # ------------------------------------------------------------------ the step
    def _eager_step(self, xb, yb, lr_scale: float = 1.0):
        self.zero_grad()
        loss, _ = self.forward_backward(xb, yb)
        self.optimizer_step(lr_scale)
        return loss

    def capture(self, B: int, T: int, warmup: int = 3, allow_collectives: bool = False):
        """Capture forward + backward (+ bucketed all-reduce) + AdamW for a fixed (B, T) into one CUDA graph
        : removes every launch gap and all Python work from the step.  step() then replays it.  Dropout works
        under the graph: the Philox seed / base offset live in device memory (ops.DeviceGenerator) and the captured
        step ends with the kernel that advances the base, so every replay draws new masks.
        With more than one rank the NCCL all-reduces and the side stream they run on are part of the graph
        (allow_collectives=True; every rank must capture, and replay, in lock step)."""
        if self.world > 1 and not allow_collectives:
            raise RuntimeError("capture(): pass allow_collectives=True to capture the bucketed all-reduce as well")
        self._gx = torch.zeros((B, T), dtype=torch.int64, device=self._dev)
        self._gy = torch.zeros((B, T), dtype=torch.int64, device=self._dev)
        self._gx[:, :] = 4
        self._gy[:, :-1] = 4
        snap = [(g.flat.clone(), g.m.clone(), g.v.clone()) for g in self.groups]
        gen_snap = None if self._gen is None else self._gen.state.clone()
        count = self.step_count
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step(self._gx, self._gy)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        self.step_count += 1
        self._push_hyper(1.0)
        self.step_count -= 1
        with torch.cuda.graph(graph):
            self.zero_grad()
            loss, _ = self.forward_backward(self._gx, self._gy)
            for bk in self.buckets:
                bk.finish()
            gscale = 1.0 / self.world
            for gi, g in enumerate(self.groups):
                ops.adamw(g.flat, self._grad_source(gi), g.m, g.v, g.shadow, g.lr, self.betas[0], self.betas[1], self.eps,
                          g.weight_decay, 1, gscale, dev_hyper=self._hyper[gi])
            self._gloss = loss
        # undo the warm-up / capture-time updates: capture must not change the training state
        for g, (f, m, v) in zip(self.groups, snap):
            g.flat.copy_(f)
            g.m.copy_(m)
            g.v.copy_(v)
            g.shadow.copy_(f)
        if gen_snap is not None:
            self._gen.state.copy_(gen_snap)
        self.step_count = count
        bump_shadow_generation()
        self._graph = graph
        self._graph_shape = (B, T)

    def step(self, xb, yb, lr_scale: float = 1.0):
        """Device-resident inputs; returns the total loss as a 0-dim device tensor (no host sync).  Under a captured
        graph the returned tensor is the graph's OWN output buffer: every replay overwrites it, so a caller that
        collects losses over several steps must `.clone()` (or `.item()`) each one before the next step."""
        if self._graph is not None and tuple(xb.shape) == self._graph_shape:
            self._gx.copy_(xb, non_blocking=True)
            self._gy.copy_(yb, non_blocking=True)
            self.step_count += 1
            self._push_hyper(lr_scale)
            self._graph.replay()
            bump_shadow_generation()
            return self._gloss
        return self._eager_step(xb, yb, lr_scale)

    def step_host(self, xb_pinned, yb_pinned, lr_scale: float = 1.0) -> float:
        """Host buffers in, host scalar out: H2D of the batch and D2H of the loss are part of the call."""
        if self._xb is None or self._xb.shape != xb_pinned.shape:
            self._xb = torch.empty(xb_pinned.shape, dtype=torch.int64, device=self._dev)
            self._yb = torch.empty(yb_pinned.shape, dtype=torch.int64, device=self._dev)
        self._xb.copy_(xb_pinned, non_blocking=True)
        self._yb.copy_(yb_pinned, non_blocking=True)
        return float(self.step(self._xb, self._yb, lr_scale).item())


# ----------------------------------------------------------------------------------------------------------
# The caller's loop around the step: learning-rate schedule, accumulation groups, non-finite policy
# ----------------------------------------------------------------------------------------------------------
class NonfiniteGroupLimitError(RuntimeError):
    """More accumulation groups were aborted than `max_nonfinite_accumulation_groups` allows (loop.py:66-67)."""


def resolve_warmup_steps(cfg: dict, total_steps: int) -> int:
    """`warmup_steps` (default 200) or `warmup_fraction` of the schedule, never both (loop.py:70-87)."""
    if total_steps <= 0:
        raise ValueError("scheduler_total_steps must be positive")
    frac = cfg.get("warmup_fraction")
    if frac is not None:
        if "warmup_steps" in cfg:
            raise ValueError("configure only one of warmup_steps or warmup_fraction")
        frac = float(frac)
        if not 0.0 <= frac < 1.0:
            raise ValueError("warmup_fraction must be in [0, 1)")
        return 0 if frac == 0.0 else max(1, int(round(total_steps * frac)))
    steps = int(cfg.get("warmup_steps", 200))
    if steps < 0:
        raise ValueError("warmup_steps must be non-negative")
    return steps


def cosine_lr_scale(step_idx: int, warmup_steps: int, total_steps: int, min_lr_ratio: float) -> float:
    """Multiplier of the base learning rates at optimiser step `step_idx` (0-based): linear warm-up then cosine
    decay to min_lr/base_lr — the reference's LambdaLR lambda (loop.py:772-778)."""
    w = max(1, int(warmup_steps))
    if step_idx < w:
        return float(step_idx + 1) / w
    progress = (step_idx - w) / max(1, total_steps - w)
    return min_lr_ratio + (1.0 - min_lr_ratio) * 0.5 * (1.0 + math.cos(math.pi * progress))


class AccumulationHealth:
    """Checkpointable counters of gradient-accumulation group integrity (loop.py:90-141): same fields, same
    state_dict, so a resumed run continues the reference's bookkeeping."""

    FIELDS = ("active_microbatches", "nonfinite_microbatches", "aborted_groups", "discarded_finite_microbatches")

    def __init__(self):
        self.active_microbatches = 0
        self.nonfinite_microbatches = 0
        self.aborted_groups = 0
        self.discarded_finite_microbatches = 0

    def record_finite_microbatch(self):
        self.active_microbatches += 1

    def complete_group(self):
        if self.active_microbatches <= 0:
            raise ValueError("cannot complete an empty accumulation group")
        self.active_microbatches = 0

    def abort_group(self) -> int:
        discarded = self.active_microbatches
        self.nonfinite_microbatches += 1
        self.aborted_groups += 1
        self.discarded_finite_microbatches += discarded
        self.active_microbatches = 0
        return discarded

    def exceeds_limit(self, max_aborted_groups: int) -> bool:
        return max_aborted_groups >= 0 and self.aborted_groups > max_aborted_groups

    def metrics_dict(self) -> Dict[str, int]:
        return {k: int(getattr(self, k)) for k in self.FIELDS}

    def state_dict(self) -> Dict[str, int]:
        state = self.metrics_dict()
        state["active_microbatches"] = 0  # gradients are not checkpointed: resume replays from the last resolved group
        return state

    def load_state_dict(self, state: Optional[dict]):
        state = state or {}
        self.active_microbatches = 0
        for k in self.FIELDS[1:]:
            setattr(self, k, int(state.get(k, 0)))


def run_accumulation_groups(step, microbatches: Iterable, grad_accum_steps: int, health: AccumulationHealth,
                            max_nonfinite_groups: int = 3, lr_scale_fn=None, first_step_idx: int = 0,
                            process_group=None):
    """Drive `step` (a TrainStep) over an iterable of (xb, yb) micro-batches with the reference's accumulation-group
    semantics (loop.py:1195-1262): gradients of `grad_accum_steps` finite micro-batches are summed and averaged
    into ONE optimiser step; a micro-batch whose loss is not finite aborts the group — its gradients and those of
    the finite micro-batches before it are discarded, the next micro-batch opens a new group — and more than
    `max_nonfinite_groups` aborts raise NonfiniteGroupLimitError; a trailing partial group still steps.

    The reference decides this with a host sync per micro-batch.  Here a group is launched back to back and
    its per-micro-batch finite flags are read ONCE (one D2H of `grad_accum_steps` floats, max-reduced over the
    data-parallel ranks so that every rank takes the same decision); only a group that did contain a non-finite
    loss is replayed from the micro-batch after it, which reproduces the reference's group boundaries exactly.

    Yields one dict per optimiser step: {"step", "group_size", "total_loss_sum", "next_loss_sum", "lr_scale"}."""
    gacc = max(1, int(grad_accum_steps))
    step_idx = int(first_step_idx)
    distributed = process_group is not None or (dist.is_available() and dist.is_initialized()
                                                and dist.get_world_size() > 1)
    queue: List[tuple] = []
    it = iter(microbatches)
    exhausted = False
    while True:
        while len(queue) < gacc and not exhausted:
            try:
                queue.append(next(it))
            except StopIteration:
                exhausted = True
        group = queue[:gacc]
        # The ranks agree on the composition of the group BEFORE running it: a rank-sharded stream whose length is not
        # a multiple of gacc x world leaves the ranks with different numbers of local micro-batches in the trailing
        # group (possibly none).  n_global (sum over ranks) is what the gradient mean divides by on EVERY rank, and a
        # rank without local micro-batches still joins the group's collectives (with zero gradients).
        n_local = len(group)
        n_global = n_local
        dev = getattr(step, "_dev", None) or torch.device("cpu")
        if distributed:
            cnt = torch.tensor([float(n_local)], device=dev)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=process_group)
            n_global = int(round(cnt.item()))
        if n_global == 0:
            return
        step.zero_grad()
        losses, nexts = [], []
        for k, (xb, yb) in enumerate(group):
            step.arm_collectives(k == n_local - 1)
            loss, parts = step.forward_backward(xb, yb)
            losses.append(loss.detach().reshape(1).float())
            nexts.append(parts["next"].detach().reshape(1).float())
        if n_local == 0:
            step.arm_collectives(True)
            join = getattr(step, "join_collectives_without_backward", None)
            if join is not None:
                join()  # the peers' hooks launch their buckets during backward, BEFORE the flag all-reduce below
        # fixed-length vectors (gacc slots; empty slots are finite zeros) so that every rank reduces the same shape
        pad = [torch.zeros(1, device=dev)] * (gacc - n_local)
        stacked = torch.cat([t.to(dev) for t in losses] + pad + [t.to(dev) for t in nexts] + pad)
        bad = (~torch.isfinite(stacked[:gacc])).float()
        if distributed:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=process_group)
        host = torch.cat([bad, stacked]).cpu()  # the group's only host sync
        bad_h = host[:gacc].tolist()
        first_bad = next((k for k, b in enumerate(bad_h) if b > 0), None)
        if first_bad is None:
            for _ in group:
                health.record_finite_microbatch()
            scale = 1.0 if lr_scale_fn is None else float(lr_scale_fn(step_idx))
            if distributed:
                step.optimizer_step(lr_scale=scale, micro_batches=max(1, n_local), global_micro_batches=n_global)
            else:
                step.optimizer_step(lr_scale=scale, micro_batches=n_local)
            if n_local:
                health.complete_group()
            vals = host[gacc:].tolist()
            yield {"step": step_idx, "group_size": n_local, "global_group_size": n_global,
                   "total_loss_sum": sum(vals[:n_local]), "next_loss_sum": sum(vals[gacc:gacc + n_local]),
                   "lr_scale": scale}
            step_idx += 1
            queue = queue[n_local:]
        else:
            for _ in range(min(first_bad, n_local)):
                health.record_finite_microbatch()
            health.abort_group()
            step.discard_gradients()
            queue = queue[first_bad + 1:]
            if health.exceeds_limit(max_nonfinite_groups):
                raise NonfiniteGroupLimitError(
                    f"nonfinite accumulation groups exceeded configured maximum {max_nonfinite_groups}: "
                    f"{health.aborted_groups}")
