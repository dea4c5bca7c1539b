"""Golden trajectory of the UNMODIFIED reference trainer (src/codonlm/training/loop.py `run_training`) on CPU:

  * a tiny config (2L2H d32, seq 32, batch 4, grad_accum 2, 44 training sequences -> 11 micro-batches per epoch: five
    full accumulation groups + one trailing partial group, cosine schedule with warm-up), 3 epochs;
  * every training micro-batch's loss and a checksum of its tokens (observed by wrapping TinyGPT.forward — the trainer
    itself is untouched), the per-epoch train / val losses, the payload keys and counters of its `last.pt`;
  * the same run INTERRUPTED at the first micro-batch of epoch 3 (the observer raises) and RESUMED by the reference
    from the `last.pt` it wrote after epoch 2: that checkpoint is the resume fixture, the resumed epoch-3 trajectory
    the golden for resuming it.

Outputs: tests/golden/trainer_golden.json, tests/golden/trainer_data.npz (the NPZ inputs) and
tests/golden/trainer_ref_last.pt (the reference's checkpoint after epoch 2: the resume fixture).

    python tests/golden/make_trainer_golden.py      (build container only: needs /root/reference)
"""
import json
import os
import shutil
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import yaml

REF = os.environ.get("CGPT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from src.codonlm.model_tiny_gpt import TinyGPT  # noqa: E402
from src.codonlm.training.loop import run_training  # noqa: E402

HERE = Path(os.path.dirname(os.path.abspath(__file__)))

CFG = {"vocab_size": 68, "block_size": 32, "n_layer": 2, "n_head": 2, "n_embd": 32, "dropout": 0.0, "batch_size": 4,
       "grad_accum_steps": 2, "max_nonfinite_accumulation_groups": 3, "lr": 0.003, "min_lr": 0.0003,
       "weight_decay": 0.05, "warmup_steps": 2, "epochs": 3, "optimizer": "adamw", "amp": False,
       "use_checkpoint": False, "scheduler": "cosine", "early_stop_patience": 5, "seed": 7, "num_workers": 0,
       "use_sdpa": True, "label_smoothing": 0.05, "device": "cpu"}


def make_data(seed=3, n_train=44, n_val=8, T=32):
    """Sequences with a learnable structure: BOS, then a motif of period 5 drawn per sequence with 10 % noise, EOS."""
    rng = np.random.default_rng(seed)
    def one():
        motif = rng.integers(4, 68, size=5)
        seq = np.array([1] + [int(motif[i % 5]) if rng.random() > 0.1 else int(rng.integers(4, 68)) for i in range(T)])
        n = int(rng.integers(T // 2, T + 1))
        seq[n] = 2
        seq[n + 1:] = 0
        return seq[:T], np.concatenate([seq[1:T], [0]])
    tr = [one() for _ in range(n_train)]
    va = [one() for _ in range(n_val)]
    return (np.stack([a for a, _ in tr]).astype(np.int32), np.stack([b for _, b in tr]).astype(np.int32),
            np.stack([a for a, _ in va]).astype(np.int32), np.stack([b for _, b in va]).astype(np.int32))


def main():
    xt, yt, xv, yv = make_data()
    np.savez_compressed(HERE / "trainer_data.npz", X_train=xt, Y_train=yt, X_val=xv, Y_val=yv)
    out = {"cfg": CFG}
    observed = []
    original_forward = TinyGPT.forward

    def observing_forward(self, idx, targets=None, *a, **kw):
        res = original_forward(self, idx, targets, *a, **kw)
        if self.training and targets is not None:
            observed.append({"loss": float(res[1].detach()), "token_sum": int(idx.sum()), "rows": int(idx.shape[0])})
        return res

    TinyGPT.forward = observing_forward
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        os.chdir(td)
        (td / "itos.txt").write_text("\n".join(f"token_{i}" for i in range(68)) + "\n")
        cfg = dict(CFG, itos_path=str(td / "itos.txt"), out_dir=str(td / "unused-ckpt"), scores_dir=str(td / "unused-scores"))
        (td / "config.yaml").write_text(yaml.safe_dump(cfg))
        np.savez_compressed(td / "train.npz", X=xt, Y=yt)
        np.savez_compressed(td / "val.npz", X=xv, Y=yv)
        args = SimpleNamespace(config=str(td / "config.yaml"), run_id="golden", resume=None, transfer_from=None,
                               train_npz=[str(td / "train.npz")], val_npz=[str(td / "val.npz")],
                               test_npz=[str(td / "val.npz")])
        run_training(dict(cfg), args)
        out["train_microbatches"] = list(observed)
        last = td / "runs" / "golden" / "checkpoints" / "last.pt"
        ck = torch.load(last, map_location="cpu", weights_only=False)
        out["payload_keys"] = sorted(ck.keys())
        out["counters"] = {k: ck[k] for k in ("epoch", "step", "best_epoch", "no_improve", "batch_size", "grad_accum_steps",
                                               "train_examples", "train_batches", "epoch_microbatch_idx")}
        out["losses"] = {k: float(ck[k]) for k in ("train_loss", "val_loss", "train_next_loss", "val_next_loss", "best_val")}
        out["scheduler"] = {k: ck["scheduler"][k] for k in ("last_epoch", "_step_count", "_last_lr", "base_lrs")}
        out["optimizer_groups"] = [{k: (list(v) if isinstance(v, tuple) else v) for k, v in pg.items() if k != "params"}
                                   | {"n_params": len(pg["params"])} for pg in ck["optimizer"]["param_groups"]]
        out["param_norms"] = {k: float(v.float().norm()) for k, v in ck["model"].items() if v.dtype.is_floating_point
                              and not k.endswith("attn.mask")}
        out["run_tree"] = sorted(str(p.relative_to(td / "runs" / "golden")) for p in (td / "runs" / "golden").rglob("*"))
        out["curves"] = {p.name: p.read_text() for p in (td / "runs" / "golden").rglob("*.csv")}

        # second run, same config, new run id: interrupted at the first micro-batch of epoch 3, then resumed
        class Interrupt(Exception):
            pass
        observed.clear()
        state = {"armed": True}

        def interrupting_forward(self, idx, targets=None, *a, **kw):
            if self.training and targets is not None and state["armed"] and len(observed) == 22:
                raise Interrupt()
            return observing_forward(self, idx, targets, *a, **kw)

        TinyGPT.forward = interrupting_forward
        args_b = SimpleNamespace(**{**vars(args), "run_id": "golden-b"})
        try:
            run_training(dict(cfg), args_b)
            raise SystemExit("the interrupting observer did not fire")
        except Interrupt:
            pass
        last_b = td / "runs" / "golden-b" / "checkpoints" / "last.pt"
        ck2 = torch.load(last_b, map_location="cpu", weights_only=False)
        out["fixture_counters"] = {k: ck2[k] for k in ("epoch", "step")}
        assert [o["loss"] for o in observed] == [o["loss"] for o in out["train_microbatches"][:22]]
        shutil.copy(last_b, HERE / "trainer_ref_last.pt")
        observed.clear()
        state["armed"] = False
        run_training(dict(cfg), SimpleNamespace(**{**vars(args_b), "resume": str(last_b)}))
        out["resumed_microbatches"] = list(observed)
        ck3 = torch.load(last_b, map_location="cpu", weights_only=False)
        out["resumed_counters"] = {k: ck3[k] for k in ("epoch", "step")}
        out["resumed_losses"] = {k: float(ck3[k]) for k in ("train_loss", "val_loss")}
    TinyGPT.forward = original_forward
    os.chdir(HERE)
    with open(HERE / "trainer_golden.json", "w") as f:
        json.dump(out, f, indent=1, default=str)
    print("train losses:", [round(o["loss"], 4) for o in out["train_microbatches"]])
    print("resumed     :", [round(o["loss"], 4) for o in out["resumed_microbatches"]])
    print(out["counters"], out["losses"], os.path.getsize(HERE / "trainer_ref_last.pt") // 1024, "KiB")


if __name__ == "__main__":
    main()
