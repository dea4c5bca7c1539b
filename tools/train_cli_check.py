"""Run `python -m codonlm_b200.train` (the reference CLI's arguments) on the trainer golden's data: writes the NPZ inputs
and a YAML config, trains on 1 GPU and — when more are visible — on 2 GPUs through the self-spawned data-parallel path,
checks the run directories and that both runs reach the same step count and a close final validation loss.
Usage (GPU box): python tools/train_cli_check.py [workdir]"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "genomics-lm_b200")


def main():
    work = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="cgpt_train_"))
    os.makedirs(work, exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "trainer_golden.json")) as f:
        g = json.load(f)
    z = np.load(os.path.join(ROOT, "tests", "golden", "trainer_data.npz"))
    np.savez_compressed(os.path.join(work, "train.npz"), X=z["X_train"], Y=z["Y_train"])
    np.savez_compressed(os.path.join(work, "val.npz"), X=z["X_val"], Y=z["Y_val"])
    cfg = dict(g["cfg"])
    cfg.pop("device", None)
    cfg["runs_dir"] = os.path.join(work, "runs")
    cfg["collective_timeout_s"] = 60  # a mismatched collective must fail the check quickly
    with open(os.path.join(work, "config.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    env = dict(os.environ, PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""))
    results = {}
    for gpus in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        run_id = f"cli_{gpus}gpu"
        cmd = [sys.executable, "-m", "codonlm_b200.train", "--config", os.path.join(work, "config.yaml"), "--run_id", run_id,
               "--train_npz", os.path.join(work, "train.npz"), "--val_npz", os.path.join(work, "val.npz"), "--gpus", str(gpus)]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=work, timeout=300)
        print(res.stdout[-1500:])
        if res.returncode != 0:
            print(res.stderr[-3000:])
            raise SystemExit(f"train CLI failed with {gpus} GPU(s)")
        run = os.path.join(work, "runs", run_id)
        for rel in ("checkpoints/last.pt", "checkpoints/best.pt", "checkpoints/meta.json", "scores/curves.csv"):
            assert os.path.exists(os.path.join(run, rel)), rel
        ck = torch.load(os.path.join(run, "checkpoints", "last.pt"), map_location="cpu", weights_only=False)
        results[gpus] = (ck["step"], ck["epoch"], float(ck["val_loss"]), float(ck["train_loss"]))
        print(f"[train_cli_check] {gpus} GPU(s): step {ck['step']} epoch {ck['epoch']} train {ck['train_loss']:.4f} "
              f"val {ck['val_loss']:.4f} (reference trainer: train {g['losses']['train_loss']:.4f} val {g['losses']['val_loss']:.4f})")
    assert results[1][0] == g["counters"]["step"]
    assert abs(results[1][2] / g["losses"]["val_loss"] - 1) < 5e-3
    if 2 in results:
        # two ranks, grad_accum 2: every optimiser step consumes the same two micro-batches, one per rank; the trailing
        # single micro-batch of an epoch is held by rank 0 alone (ragged group, agreed on by both ranks)
        assert results[2][0] == results[1][0], results
        assert abs(results[2][2] / results[1][2] - 1) < 1e-2, results
        assert abs(results[2][3] / results[1][3] - 1) < 1e-2, results  # epoch mean over BOTH ranks' micro-batches
    print("[train_cli_check] ok")


if __name__ == "__main__":
    main()
