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

    def optimizer_step(self, lr_scale=1.0, micro_batches=1):
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
