"""World-size-2 gloo test (CPU) of the data-parallel host logic: flat parameter groups, bucket
partitioning, gradient-ready hooks and the bucketed bf16 all-reduce.  The reduced gradient must equal the
sum of the two ranks' gradients (to bf16 rounding) and both ranks must end up with identical buffers."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG


def _worker(rank, world, port, out):
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from codonlm_b200.trainer import FlatGroup, GradBuckets
    torch.manual_seed(0)  # same weights on both ranks
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                              torch.nn.Linear(64, 8))
    named = list(net.named_parameters())
    named.reverse()  # reverse execution order, as split_param_groups does
    group = FlatGroup(named, lr=1e-3, weight_decay=0.0)
    buckets = GradBuckets(group, dist.group.WORLD, bucket_bytes=1024)  # several buckets
    assert len(buckets.bounds) >= 3 and buckets.bounds[0][0] == 0 and buckets.bounds[-1][1] == group.numel
    assert all(a[1] == b[0] for a, b in zip(buckets.bounds, buckets.bounds[1:]))
    assert sum(buckets.need) == len(group.params)
    results = []
    for it in range(2):  # two steps: the ready-counters must re-arm
        group.grad.zero_()
        for p_ in group.params:  # plain autograd (no main_grad-aware kernels here): accumulate into the flat views
            p_.grad = p_.main_grad
        torch.manual_seed(100 + rank + 10 * it)  # different data per rank
        x = torch.randn(32, 16)
        net(x).square().mean().backward()
        local = group.grad.clone()
        if it > 0:  # after the counting step every bucket is launched from the hooks, before finish()
            assert len(buckets.pending) == len(buckets.bounds)
        buckets.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        expect = sum(g.to(torch.bfloat16).float() for g in gathered)
        err = (group.grad - expect).abs().max().item()
        results.append((err, expect.abs().max().item()))
        both = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(both, group.grad)
        assert torch.equal(both[0], both[1])
    if rank == 0:
        torch.save(results, out)
    dist.barrier()  # no rank tears its gloo threads down while the other one still talks to it
    dist.destroy_process_group()


def test_bucketed_allreduce_world2(tmp_path):
    out = str(tmp_path / "res.pt")
    import socket
    with socket.socket() as sk:  # a port that is free right now (a fixed formula collided with other runs)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    try:
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    except mp.ProcessExitedException:
        # gloo's background threads occasionally abort a worker while it EXITS ("terminate called without an active
        # exception"); every assertion of the worker ran before rank 0 wrote the results, so they still count
        if not os.path.exists(out):
            raise
    for err, scale in torch.load(out):
        assert err <= 2 ** -7 * scale  # bf16 all-reduce: two roundings


def test_param_groups_follow_reference_rule():
    from codonlm_b200 import TinyGPT
    from codonlm_b200.trainer import split_param_groups
    m = TinyGPT(68, 16, n_layer=2, n_head=2, n_embd=32, dropout=0.0, termination_aux=True, multi_offset_targets=[2, 4])
    g = split_param_groups(m)
    head_names = [n for n, _ in g["head"]]
    assert head_names and all(("offset_projs" in n) or ("termination_head" in n) for n in head_names)
    back = [n for n, _ in g["backbone"]]
    assert "tok_emb.weight" in back and "ln_f.bias" in back and "head.weight" not in back  # tied: listed once
    assert back[-1] == "tok_emb.weight" and back[0].startswith("ln_f")  # reverse execution order
    n_total = sum(p.numel() for p in m.parameters())
    assert sum(p.numel() for _, p in g["head"] + g["backbone"]) == n_total


def test_flat_buffers_keep_packed_operands_adjacent():
    """FlatGroup on the trainer's parameter order: the query|key|value weights (and biases) of every block must lie
    back to back in the flat master / gradient / shadow buffers — that is what lets one GEMM, one weight-gradient GEMM
    and the attention backward's fused bias sums treat them as ONE packed operand (functional._packed_slots)."""
    from codonlm_b200 import TinyGPT
    from codonlm_b200.functional import _packed_slots
    from codonlm_b200.model_tiny_gpt import _adjacent
    from codonlm_b200.trainer import FlatGroup, split_param_groups
    for kw in (dict(), dict(n_kv_head=2), dict(use_swiglu=True, use_rope=True)):
        m = TinyGPT(68, 32, n_layer=3, n_head=4, n_embd=64, dropout=0.0, termination_aux=True,
                    multi_offset_targets=[2, 4], **kw)
        g = split_param_groups(m)
        fg = FlatGroup(g["backbone"], lr=1e-3, weight_decay=0.0)
        assert fg.flat.numel() == fg.grad.numel() == fg.shadow.numel()
        for blk in m.blocks:
            a = blk.attn
            ws = (a.query.weight, a.key.weight, a.value.weight)
            bs = (a.query.bias, a.key.bias, a.value.bias)
            assert _adjacent([p.data for p in ws]) and _adjacent([p.main_grad for p in ws])
            assert _adjacent([p._cgpt_shadow for p in ws])
            rows, r = [], 0
            for p in ws:
                rows.append((r, p.shape[0]))
                r += p.shape[0]
            assert _packed_slots(ws, rows) is not None
            assert _packed_slots((a.key.weight, a.query.weight, a.value.weight), rows) is None  # wrong order
            # every slot starts on a 256-byte boundary: a bias shorter than 64 floats (GQA with a narrow kv width) leaves
            # a gap, and the module then falls back to per-bias column sums / a private packed copy
            tight = all(p.numel() % 64 == 0 for p in bs)
            assert _adjacent([p.data for p in bs]) == tight
            assert (_packed_slots(bs, rows) is not None) == tight
        # every parameter is a view of the flat buffer and keeps its values
        for n, p in g["backbone"]:
            assert p.data.data_ptr() >= fg.flat.data_ptr() and p.main_grad.shape == p.shape


def _nonfinite_worker(rank, world, port, out):
    import math
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from codonlm_b200.trainer import AccumulationHealth, run_accumulation_groups

    class FakeStep:
        def __init__(self):
            self.grad, self.steps, self.discards = 0.0, [], 0

        def zero_grad(self):
            self.grad = 0.0

        def arm_collectives(self, armed):
            pass

        def forward_backward(self, xb, yb):
            self.grad += 0.0 if math.isnan(xb) else xb
            loss = torch.tensor(float(xb))
            return loss, {"next": loss}

        def optimizer_step(self, lr_scale=1.0, micro_batches=1, global_micro_batches=None):
            self.steps.append((self.grad, micro_batches))

        def discard_gradients(self):
            self.discards += 1
            self.grad = 0.0

    # only rank 1 sees a non-finite loss (its 4th micro-batch): BOTH ranks must abort that group and regroup alike
    vals = [1.0, 2.0, 3.0, float("nan") if rank == 1 else 4.0, 5.0, 6.0, 7.0]
    step, health = FakeStep(), AccumulationHealth()
    recs = list(run_accumulation_groups(step, [(v, None) for v in vals], 2, health, max_nonfinite_groups=3))
    torch.save(dict(steps=step.steps, discards=step.discards, health=health.metrics_dict(),
                    sizes=[r["group_size"] for r in recs]), f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_nonfinite_flag_is_shared_across_ranks(tmp_path):
    """run_accumulation_groups max-reduces the per-micro-batch non-finite flags over the ranks: a NaN seen
    by one rank aborts the accumulation group on every rank, so the replicas keep taking the same optimiser steps."""
    import socket
    out = str(tmp_path / "nf.pt")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    try:
        mp.spawn(_nonfinite_worker, args=(2, port, out), nprocs=2, join=True)
    except mp.ProcessExitedException:
        if not (os.path.exists(out + ".0") and os.path.exists(out + ".1")):
            raise
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    # groups: [1,2] step | [3,(4|nan)] aborted on both | [5,6] step | [7] trailing step
    assert r0["sizes"] == r1["sizes"] == [2, 2, 1]
    assert r0["discards"] == r1["discards"] == 1
    assert r0["health"] == r1["health"] == {"active_microbatches": 0, "nonfinite_microbatches": 1, "aborted_groups": 1,
                                            "discarded_finite_microbatches": 1}
    assert [s[0] for s in r0["steps"]] == [3.0, 11.0, 7.0] and [s[0] for s in r1["steps"]] == [3.0, 11.0, 7.0]


def _ragged_worker(rank, world, port, out):
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from codonlm_b200.token_feed import rank_microbatches
    from codonlm_b200.trainer import AccumulationHealth, run_accumulation_groups

    class SumStep:
        """'Gradient' = sum of local micro-batch values; optimizer_step all-reduces it like the gradient buckets do and
        records the mean the fused AdamW would see (sum over ranks x 1/global count)."""

        def __init__(self):
            self.grad, self.means, self.calls = torch.zeros(1), [], []

        def zero_grad(self):
            self.grad.zero_()

        def arm_collectives(self, armed):
            pass

        def forward_backward(self, xb, yb):
            self.grad += xb
            loss = torch.tensor(float(xb))
            return loss, {"next": loss}

        def optimizer_step(self, lr_scale=1.0, micro_batches=1, global_micro_batches=None):
            g = self.grad.clone()
            dist.all_reduce(g)
            self.means.append(g.item() / global_micro_batches)
            self.calls.append((micro_batches, global_micro_batches))

        def discard_gradients(self):
            self.grad.zero_()

    # 11 micro-batches, global accumulation 4, world 2: groups [0..3] [4..7] [8,9,10]; the last one gives rank 0 two
    # micro-batches (8, 10) and rank 1 one (9) — ADVICE r1: ranks must agree on the group and divide by the GLOBAL count
    vals = [float(v) for v in range(1, 12)]
    mine = list(rank_microbatches([(v, None) for v in vals], rank, world, 4))
    step, health = SumStep(), AccumulationHealth()
    recs = list(run_accumulation_groups(step, mine, 4 // world, health))
    # 9 micro-batches: the trailing group is [9] and only rank 0 holds it — rank 1 must still join its collectives
    vals2 = [float(v) for v in range(1, 10)]
    mine2 = list(rank_microbatches([(v, None) for v in vals2], rank, world, 4))
    step2 = SumStep()
    recs2 = list(run_accumulation_groups(step2, mine2, 4 // world, AccumulationHealth()))
    torch.save(dict(means=step.means, calls=step.calls, sizes=[(r["group_size"], r["global_group_size"]) for r in recs],
                    means2=step2.means, sizes2=[(r["group_size"], r["global_group_size"]) for r in recs2]),
               f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_trailing_group_is_agreed_on_by_all_ranks(tmp_path):
    import socket
    out = str(tmp_path / "rg.pt")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    try:
        mp.spawn(_ragged_worker, args=(2, port, out), nprocs=2, join=True)
    except mp.ProcessExitedException:
        if not (os.path.exists(out + ".0") and os.path.exists(out + ".1")):
            raise
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    want = [sum(range(1, 5)) / 4, sum(range(5, 9)) / 4, sum(range(9, 12)) / 3]  # the single-process reference's means
    assert r0["means"] == pytest.approx(want) and r1["means"] == pytest.approx(want)
    assert r0["sizes"] == [(2, 4), (2, 4), (2, 3)] and r1["sizes"] == [(2, 4), (2, 4), (1, 3)]
    want2 = [sum(range(1, 5)) / 4, sum(range(5, 9)) / 4, 9.0]
    assert r0["means2"] == pytest.approx(want2) and r1["means2"] == pytest.approx(want2)
    assert r0["sizes2"][-1] == (1, 1) and r1["sizes2"][-1] == (0, 1)


def _empty_rank_worker(rank, world, port, out):
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import datetime
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=60))
    from codonlm_b200.trainer import FlatGroup, GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 48), torch.nn.GELU(),
                              torch.nn.Linear(48, 8))
    named = list(net.named_parameters())
    named.reverse()
    group = FlatGroup(named, lr=1e-3, weight_decay=0.0)
    buckets = GradBuckets(group, dist.group.WORLD, bucket_bytes=1024)  # several buckets of DIFFERENT sizes
    assert len({e - s for s, e in buckets.bounds}) > 1

    def backward(seed):
        group.grad.zero_()
        for p_ in group.params:
            p_.grad = p_.main_grad
        torch.manual_seed(seed)
        net(torch.randn(32, 16)).square().mean().backward()

    flags = []
    for it in range(3):  # step 0 counts contributions, step 1 launches from the hooks on both ranks
        if it < 2 or rank == 0:
            backward(100 + rank + 10 * it)
        else:  # step 2: rank 1 holds no micro-batch of this (ragged) group
            group.grad.zero_()
            buckets.join_without_backward()
        flag = torch.tensor([float(rank == 0 and it == 2)])
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)  # what run_accumulation_groups reduces right after backward
        flags.append(flag.item())
        local = group.grad.clone()
        buckets.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        expect = sum(g.to(torch.bfloat16).float() for g in gathered)
        assert (group.grad - expect).abs().max().item() <= 2 ** -7 * max(1e-6, expect.abs().max().item())
    if rank == 0:
        torch.save(flags, out)
    dist.barrier()
    dist.destroy_process_group()


def test_rank_without_microbatch_issues_collectives_in_peer_order(tmp_path):
    """A rank that holds no micro-batch of a ragged accumulation group must issue the bucket all-reduces where its
    peers' gradient-ready hooks do — before the flag all-reduce that follows backward — or the collectives pair up
    wrongly (found on 2 GPUs: the trainer CLI dead-locked in the trailing group of an epoch)."""
    import socket
    out = str(tmp_path / "er.pt")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    try:
        mp.spawn(_empty_rank_worker, args=(2, port, out), nprocs=2, join=True)
    except mp.ProcessExitedException:
        if not os.path.exists(out):
            raise
    assert torch.load(out) == [0.0, 0.0, 1.0]
