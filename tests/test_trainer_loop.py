"""CPU tests of the caller-side loop logic: LR schedule, warm-up resolution, AccumulationHealth
and the sync-free accumulation-group driver, against a plain restatement of the reference's per-micro-batch loop
(src/codonlm/training/loop.py:70-141, 772-778, 1195-1262; reference tests: tests/test_nonfinite_accumulation.py,
tests/test_warmup_schedule.py)."""
import math

import pytest
import torch

from codonlm_b200.trainer import (AccumulationHealth, NonfiniteGroupLimitError, cosine_lr_scale, resolve_warmup_steps,
                                  run_accumulation_groups)


class FakeStep:
    """Stands in for TrainStep: 'gradient' = sum of the micro-batch values, optimiser step records their mean."""

    def __init__(self):
        self.grad = 0.0
        self.steps = []
        self.armed_log = []
        self.discards = 0

    def zero_grad(self):
        self.grad = 0.0

    def arm_collectives(self, armed):
        self.armed_log.append(bool(armed))

    def forward_backward(self, xb, yb):
        self.grad += float(xb)
        loss = torch.tensor(float(xb))
        return loss, {"next": loss * 0.5}

    def optimizer_step(self, lr_scale=1.0, micro_batches=1, global_micro_batches=None):
        self.steps.append((self.grad / micro_batches, micro_batches, lr_scale))

    def discard_gradients(self):
        self.discards += 1
        self.grad = 0.0


def reference_loop(values, gacc, max_groups):
    """The reference's control flow with its host check per micro-batch."""
    steps, active = [], []
    h = dict(nonfinite_microbatches=0, aborted_groups=0, discarded_finite_microbatches=0)
    for v in values:
        if not math.isfinite(v):
            h["discarded_finite_microbatches"] += len(active)
            h["nonfinite_microbatches"] += 1
            h["aborted_groups"] += 1
            active = []
            if max_groups >= 0 and h["aborted_groups"] > max_groups:
                raise NonfiniteGroupLimitError("limit")
            continue
        active.append(v)
        if len(active) == gacc:
            steps.append((sum(active) / len(active), len(active)))
            active = []
    if active:
        steps.append((sum(active) / len(active), len(active)))
    return steps, h


@pytest.mark.parametrize("gacc", [1, 3, 4])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_groups_match_reference_control_flow(gacc, seed):
    g = torch.Generator().manual_seed(seed)
    vals = torch.rand(23, generator=g).tolist()
    for pos in torch.randperm(23, generator=g)[: seed].tolist():  # seed 0: no non-finite micro-batch at all
        vals[pos] = float("nan") if pos % 2 else float("inf")
    want_steps, want_h = reference_loop(vals, gacc, max_groups=-1)
    fake, health = FakeStep(), AccumulationHealth()
    out = list(run_accumulation_groups(fake, [(v, None) for v in vals], gacc, health, max_nonfinite_groups=-1,
                                       lr_scale_fn=lambda s: 1.0 / (s + 1)))
    assert [(round(a, 6), n) for a, n, _ in fake.steps] == [(round(a, 6), n) for a, n in want_steps]
    assert [o["step"] for o in out] == list(range(len(want_steps)))
    assert [s[2] for s in fake.steps] == [1.0 / (i + 1) for i in range(len(want_steps))]
    m = health.metrics_dict()
    assert m["active_microbatches"] == 0
    for k, v in want_h.items():
        assert m[k] == v, k
    assert fake.discards == want_h["aborted_groups"]
    for o, (mean, n) in zip(out, want_steps):
        assert o["group_size"] == n and o["total_loss_sum"] == pytest.approx(mean * n)
        assert o["next_loss_sum"] == pytest.approx(0.5 * mean * n)


def test_only_last_microbatch_of_a_group_arms_the_collectives():
    fake = FakeStep()
    list(run_accumulation_groups(fake, [(1.0, None)] * 7, 3, AccumulationHealth()))
    assert fake.armed_log == [False, False, True, False, False, True, True]  # trailing partial group of one


def test_abort_limit_raises_like_reference():
    vals = [1.0, float("nan"), 1.0, float("nan"), 2.0, float("nan"), 3.0]
    with pytest.raises(NonfiniteGroupLimitError):
        list(run_accumulation_groups(FakeStep(), [(v, None) for v in vals], 2, AccumulationHealth(),
                                     max_nonfinite_groups=2))
    health = AccumulationHealth()
    list(run_accumulation_groups(FakeStep(), [(v, None) for v in vals], 2, health, max_nonfinite_groups=3))
    assert health.aborted_groups == 3 and health.discarded_finite_microbatches == 3


def test_health_state_dict_roundtrip():
    h = AccumulationHealth()
    h.record_finite_microbatch()
    h.record_finite_microbatch()
    assert h.abort_group() == 2
    h.record_finite_microbatch()
    st = h.state_dict()
    assert st == {"active_microbatches": 0, "nonfinite_microbatches": 1, "aborted_groups": 1,
                  "discarded_finite_microbatches": 2}
    h2 = AccumulationHealth()
    h2.load_state_dict(st)
    assert h2.metrics_dict() == st
    with pytest.raises(ValueError):
        h2.complete_group()
    assert not h2.exceeds_limit(-1) and h2.exceeds_limit(0) and not h2.exceeds_limit(1)


def test_cosine_schedule_matches_lambda_lr():
    base_lr, min_lr, warm, total = 3e-4, 1e-5, 5, 40
    ratio = min_lr / base_lr
    w = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([w], lr=base_lr)

    def lr_lambda(i):  # the formula of loop.py:772-778
        if i < max(1, warm):
            return float(i + 1) / max(1, warm)
        prog = (i - max(1, warm)) / max(1, total - max(1, warm))
        return ratio + (1 - ratio) * 0.5 * (1.0 + math.cos(math.pi * prog))

    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda)
    for i in range(total + 3):
        assert opt.param_groups[0]["lr"] == pytest.approx(base_lr * cosine_lr_scale(i, warm, total, ratio), rel=1e-12)
        opt.step()
        sched.step()
    assert cosine_lr_scale(0, 0, 10, 0.1) == 1.0  # warm-up of 0 behaves like 1 (max(1, warmup))


def test_resolve_warmup_steps_rules():
    assert resolve_warmup_steps({}, 1000) == 200
    assert resolve_warmup_steps({"warmup_steps": 7}, 1000) == 7
    assert resolve_warmup_steps({"warmup_fraction": 0.05}, 1000) == 50
    assert resolve_warmup_steps({"warmup_fraction": 0.0}, 1000) == 0
    assert resolve_warmup_steps({"warmup_fraction": 1e-6}, 1000) == 1
    for bad in ({"warmup_fraction": 0.1, "warmup_steps": 3}, {"warmup_fraction": 1.0}, {"warmup_steps": -1}):
        with pytest.raises(ValueError):
            resolve_warmup_steps(bad, 1000)
    with pytest.raises(ValueError):
        resolve_warmup_steps({}, 0)


# ---- optimiser state in the reference's checkpoint layout (loop.py:891-893, 962-963) -------------------------------
def _small_model(**kw):
    from codonlm_b200 import TinyGPT
    torch.manual_seed(11)
    base = dict(vocab_size=68, block_size=16, n_layer=2, n_head=2, n_embd=32, dropout=0.0, termination_aux=True,
                multi_offset_targets=[2, 4])
    base.update(kw)
    return TinyGPT(**base)


def _reference_adamw(model, lr, lr_embedding, wd):
    """The reference's optimiser construction (loop.py:681-731), restated."""
    fast, slow = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (fast if ("transformer.wte" in name or "shape_proj" in name or "offset_projs" in name
                  or "termination_head" in name) else slow).append(p)
    groups = []
    if fast:
        groups.append({"params": fast, "lr": lr_embedding, "weight_decay": 0.0})
    if slow:
        groups.append({"params": slow, "lr": lr, "weight_decay": wd})
    return torch.optim.AdamW(groups)


def test_optimizer_state_round_trips_with_torch_adamw():
    """TrainStep.state_dict() must be loadable by the reference's torch.optim.AdamW (same groups, same parameter ids)
    and carry m / v / step exactly; a state_dict written by that AdamW must load into the flat buffers."""
    import copy
    from codonlm_b200.trainer import TrainStep, cosine_lr_scale
    m_ref = _small_model()
    m_ours = copy.deepcopy(m_ref)
    opt = _reference_adamw(m_ref, 3e-3, 1e-3, 0.05)
    sched_fn = lambda i: cosine_lr_scale(i, 2, 10, 0.1)  # noqa: E731
    sched = torch.optim.lr_scheduler.LambdaLR(opt, sched_fn)
    for _ in range(3):  # three real AdamW steps on CPU with synthetic gradients
        for p in m_ref.parameters():
            p.grad = torch.randn_like(p)
        opt.step()
        sched.step()
    ts = TrainStep(m_ours, lr=3e-3, lr_embedding=1e-3, weight_decay=0.05)
    ts.load_state_dict(copy.deepcopy(opt.state_dict()))
    assert ts.step_count == 3
    named_ref = dict(m_ref.named_parameters())
    for name, p in m_ours.named_parameters():
        g, o = ts._slot(p)
        st = opt.state[named_ref[name]]
        assert torch.equal(g.m[o:o + p.numel()].view_as(p), st["exp_avg"]), name
        assert torch.equal(g.v[o:o + p.numel()].view_as(p), st["exp_avg_sq"]), name
    assert [g.lr for g in ts.groups] == [1e-3, 3e-3] and [g.weight_decay for g in ts.groups] == [0.0, 0.05]
    # and back: what TrainStep writes is what torch wrote (ids, tensors, hyper-parameters, scheduled lr)
    back = ts.state_dict(lr_scale_fn=sched_fn)
    want = opt.state_dict()
    assert [pg["params"] for pg in back["param_groups"]] == [pg["params"] for pg in want["param_groups"]]
    for a, b in zip(back["param_groups"], want["param_groups"]):
        assert set(a) == set(b)
        for k in ("weight_decay", "betas", "eps", "initial_lr", "amsgrad"):
            assert a[k] == b[k], k
        assert a["lr"] == pytest.approx(b["lr"], rel=1e-12)
    assert set(back["state"]) == set(want["state"])
    for k in want["state"]:
        for f in ("step", "exp_avg", "exp_avg_sq"):
            assert torch.equal(back["state"][k][f].float(), want["state"][k][f].float()), (k, f)
    fresh = _reference_adamw(copy.deepcopy(m_ref), 3e-3, 1e-3, 0.05)
    fresh.load_state_dict(back)  # the reference's resume path (loop.py:891-893) accepts it
    ssd = ts.scheduler_state_dict(sched_fn)
    assert set(ssd) == set(sched.state_dict()) and ssd["last_epoch"] == 3
    assert ssd["_last_lr"] == pytest.approx(sched.state_dict()["_last_lr"])
    torch.optim.lr_scheduler.LambdaLR(fresh, sched_fn).load_state_dict(ssd)
    # an untouched TrainStep writes an empty state, like a fresh torch optimiser
    assert TrainStep(_small_model()).state_dict()["state"] == {}


def test_frozen_backbone_builds_only_the_head_group():
    """freeze_backbone (loop.py:656-667): only offset_projs / termination_head train; the backbone group is empty and
    must simply not exist (the reference guards with `if backbone_params:`)."""
    from codonlm_b200.trainer import TrainStep
    m = _small_model()
    for name, p in m.named_parameters():
        p.requires_grad = ("offset_projs" in name) or ("termination_head" in name)
    ts = TrainStep(m, lr=1e-3, lr_embedding=5e-4)
    assert ts.group_kinds == ["head"] and len(ts.groups) == 1 and ts.groups[0].lr == 5e-4
    assert all(("offset_projs" in n) or ("termination_head" in n) for n in ts.groups[0].names)
    assert len(ts.state_dict()["param_groups"]) == 1
    for p in m.parameters():
        p.requires_grad = False
    with pytest.raises(ValueError):
        TrainStep(m)
