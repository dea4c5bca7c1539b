"""f1 — the trainer loop re-hosted on TrainStep (`codonlm_b200.train`), against the UNMODIFIED reference trainer.

Golden: tests/golden/make_trainer_golden.py ran the reference's `run_training` on CPU (tiny config, 3 epochs of 11
micro-batches, grad_accum 2 -> a trailing partial group every epoch, cosine schedule with warm-up), recording every
micro-batch's loss + a checksum of its tokens, the `last.pt` payload, and an interrupted + resumed third epoch.

CPU tests: the epoch order equals torch's DataLoader order, the config -> model mapping, the offset-weight rule.
GPU tests: (1) the same 3 epochs here — same micro-batches in the same order, loss trajectory within bf16 tolerance,
same counters / scheduler position, payload keys a superset of the reference's, the run directory layout;
(2) the reference's own epoch-2 `last.pt` resumed HERE for epoch 3 against the reference's resumed epoch 3;
(3) a `last.pt` written here loaded by torch.optim.AdamW + LambdaLR built as the reference builds them.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

DEV = "cuda"


def _golden():
    with open(os.path.join(GOLDEN_DIR, "trainer_golden.json")) as f:
        g = json.load(f)
    z = np.load(os.path.join(GOLDEN_DIR, "trainer_data.npz"))
    return g, z


def test_epoch_order_is_the_dataloaders():
    from torch.utils.data import DataLoader, Dataset
    from codonlm_b200.train import epoch_order

    class Idx(Dataset):
        def __len__(self):
            return 44

        def __getitem__(self, i):
            return i
    for seed in (8, 9, 1338):
        gen = torch.Generator()
        gen.manual_seed(seed)
        want = [b.tolist() for b in DataLoader(Idx(), batch_size=4, shuffle=True, generator=gen)]
        assert epoch_order(44, 4, seed) == want
    assert epoch_order(10, 4, None, shuffle=False) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]


def test_config_mapping_and_offset_weights():
    from codonlm_b200.train import build_model_from_training_cfg, normalize_offset_weights
    cfg = dict(vocab_size=68, block_size=32, n_layer=1, n_head=2, n_embd=32, dropout=0.1, multi_offset_loss_enabled=True,
               multi_offset_targets=[2, 4], termination_loss_enabled=True, termination_bucket_edges=[0, 3, 10],
               sep_mask_enabled=False, n_kv_head=1, use_swiglu=True, use_rope=True, label_smoothing=0.05)
    m = build_model_from_training_cfg(cfg)
    assert m.sep_id is None and m.multi_offset_targets == [2, 4] and m.termination_n_classes == 4
    assert m.pos_emb is None and m.n_kv_head == 1 and m.label_smoothing == 0.05
    assert build_model_from_training_cfg(dict(cfg, multi_offset_loss_enabled=False)).multi_offset_targets == []
    assert normalize_offset_weights([2, 4], None) == {2: 0.5, 4: 0.5}
    assert normalize_offset_weights([2, 4], [0.3, 0.1]) == {2: 0.3, 4: 0.1}
    assert normalize_offset_weights([2, 4], {"2": 0.7, 8: 1.0}) == {2: 0.7}


def _trainer(tmp_path, g, z, run_id, epochs=None):
    from codonlm_b200.train import StaticTokenSet, Trainer
    cfg = dict(g["cfg"])
    cfg.pop("device", None)
    if epochs is not None:
        cfg["epochs"] = epochs
    itos = tmp_path / "itos.txt"
    itos.write_text("\n".join(f"token_{i}" for i in range(68)) + "\n")
    cfg["itos_path"] = str(itos)
    train = StaticTokenSet(z["X_train"], z["Y_train"], DEV)
    val = StaticTokenSet(z["X_val"], z["Y_val"], DEV)
    return Trainer(cfg, train, val, run_root=str(tmp_path / "runs"), run_id=run_id, device=DEV, log=lambda *a, **k: None)


@pytest.mark.gpu
def test_three_epochs_follow_the_reference_trainer(tmp_path):
    g, z = _golden()
    tr = _trainer(tmp_path, g, z, "rehost")
    # same initial weights as the reference run (constructor RNG contract under manual_seed(cfg.seed))
    for k, want in g["param_norms"].items():
        pass  # (norms after training; initial equality is covered by test_module_contract's init hash)
    seen = []
    orig = tr.train_epoch

    def spying(epoch_idx, **kw):
        return orig(epoch_idx, on_microbatch=lambda ind: seen.append(int(z["X_train"][np.asarray(ind)].sum())), **kw)
    tr.train_epoch = spying
    # per-micro-batch losses: read them from the accumulation records through a wrapper around forward_backward
    losses = []
    fb = tr.step_obj.forward_backward

    def fb_spy(xb, yb, **kw):
        total, parts = fb(xb, yb, **kw)
        losses.append(total)
        return total, parts
    tr.step_obj.forward_backward = fb_spy
    out = tr.fit()
    got = [float(t) for t in losses]
    want = g["train_microbatches"]
    assert len(got) == len(want) == 33
    assert seen == [w["token_sum"] for w in want]                       # the same micro-batches, in the same order
    rel = [abs(a / w["loss"] - 1) for a, w in zip(got, want)]
    assert got[0] == pytest.approx(want[0]["loss"], rel=3e-3)            # first micro-batch: same weights, bf16 forward
    assert max(rel) <= 5e-3, f"worst micro-batch loss deviation {max(rel):.3e}"  # measured 2.7e-4
    assert out["step"] == g["counters"]["step"] == 18
    hist = out["history"]
    assert len(hist) == 3
    assert hist[-1]["train_loss"] == pytest.approx(g["losses"]["train_loss"], rel=5e-3)
    assert hist[-1]["val_loss"] == pytest.approx(g["losses"]["val_loss"], rel=5e-3)
    assert out["best_epoch"] == g["counters"]["best_epoch"]
    print(f"\n[trainer re-host] worst micro-batch loss deviation {max(rel):.2e}; final train "
          f"{hist[-1]['train_loss']:.4f} vs {g['losses']['train_loss']:.4f}, val {hist[-1]['val_loss']:.4f} vs "
          f"{g['losses']['val_loss']:.4f}")
    # run directory + payload
    run = tmp_path / "runs" / "rehost"
    for rel_path in ("checkpoints/last.pt", "checkpoints/best.pt", "checkpoints/best_epoch_003.pt", "checkpoints/meta.json",
                     "checkpoints/config.yaml", "scores/curves.csv", "itos.txt", ".run.lock"):
        assert (run / rel_path).exists(), rel_path
        assert rel_path in g["run_tree"] or rel_path == ".run.lock"
    ck = torch.load(run / "checkpoints" / "last.pt", map_location="cpu", weights_only=False)
    assert set(g["payload_keys"]) <= set(ck.keys())
    for k, v in g["counters"].items():
        assert ck[k] == v, (k, ck[k], v)
    assert ck["scheduler"]["last_epoch"] == g["scheduler"]["last_epoch"]
    assert ck["scheduler"]["_last_lr"] == pytest.approx(g["scheduler"]["_last_lr"])
    pg, want_pg = ck["optimizer"]["param_groups"][0], g["optimizer_groups"][0]
    assert len(pg["params"]) == want_pg["n_params"] and pg["lr"] == pytest.approx(want_pg["lr"])
    assert pg["initial_lr"] == want_pg["initial_lr"] and pg["weight_decay"] == want_pg["weight_decay"]
    for k, want_norm in g["param_norms"].items():
        assert float(ck["model"][k].float().norm()) == pytest.approx(want_norm, rel=2e-2), k
    header = (run / "scores" / "curves.csv").read_text().splitlines()[0]
    assert header == g["curves"]["curves.csv"].splitlines()[0]
    # (3) the checkpoint written here resumes in the reference's optimiser / scheduler objects
    from codonlm_b200.train import build_model_from_training_cfg
    ref_model = build_model_from_training_cfg(ck["cfg"])
    ref_model.load_state_dict(ck["model"], strict=True)
    params = [p for _, p in ref_model.named_parameters()]
    opt = torch.optim.AdamW([{"params": params, "lr": 0.003, "weight_decay": 0.05}])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, tr.lr_scale)
    opt.load_state_dict(ck["optimizer"])
    sched.load_state_dict(ck["scheduler"])
    assert sched.last_epoch == 18 and float(opt.state[params[0]]["step"]) == 18.0


@pytest.mark.gpu
def test_reference_checkpoint_resumes_here(tmp_path):
    """The reference's own `last.pt` after epoch 2 (tests/golden/trainer_ref_last.pt) is loaded by Trainer.resume():
    weights strict, AdamW moments / step from torch's state_dict layout, scheduler position, counters; the third epoch
    then follows the reference's resumed third epoch."""
    g, z = _golden()
    tr = _trainer(tmp_path, g, z, "resumed")
    ck = tr.resume(os.path.join(GOLDEN_DIR, "trainer_ref_last.pt"))
    assert tr.start_epoch == 2 and tr.step == 12 and tr.step_obj.step_count == 12
    name, p = next(iter(tr.model.named_parameters()))
    grp, off = tr.step_obj._slot(p)
    idx = [i for i, (n, _) in enumerate(tr.model.named_parameters()) if n == name][0]
    assert torch.equal(grp.m[off:off + p.numel()].view_as(p).cpu(), ck["optimizer"]["state"][idx]["exp_avg"].cpu())
    losses = []
    fb = tr.step_obj.forward_backward

    def fb_spy(xb, yb, **kw):
        total, parts = fb(xb, yb, **kw)
        losses.append(total)
        return total, parts
    tr.step_obj.forward_backward = fb_spy
    out = tr.fit()
    got = [float(t) for t in losses]
    want = [w["loss"] for w in g["resumed_microbatches"]]
    assert len(got) == len(want) == 11
    assert got[0] == pytest.approx(want[0], rel=3e-3)  # same weights, same optimiser state, same micro-batch
    assert max(abs(a / b - 1) for a, b in zip(got, want)) <= 5e-3
    assert out["step"] == g["resumed_counters"]["step"] == 18
    assert out["history"][-1]["val_loss"] == pytest.approx(g["resumed_losses"]["val_loss"], rel=1e-2)
