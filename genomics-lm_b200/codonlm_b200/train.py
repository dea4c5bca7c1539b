"""Re-host of the reference trainer's loop around the step on `TrainStep`.

Reference: src/codonlm/training/loop.py — model construction from the YAML keys (:559-579), freeze_backbone (:656-667),
the two AdamW groups + LambdaLR cosine (:681-784), the epoch loop with accumulation groups, validation pass, best / last
checkpoints and early stopping (:1016-1440), the checkpoint payload (:950-1007) and resume (:880-940); CLI
src/codonlm/train_codon_lm.py:36-57; run directory src/training/run_lifecycle.py:250-261.

What is kept exactly: the YAML keys, the micro-batch order of every epoch (DataLoader(shuffle=True, generator seeded
`seed + epoch`), or the BucketBatchSampler order for the dynamic format), the accumulation-group semantics, the cosine
schedule, `runs/<RUN_ID>/checkpoints/{last,best,best_epoch_XXX,epoch_N}.pt` with the reference's payload keys — the
optimiser / scheduler entries in torch's own state_dict layout, so a `last.pt` written by the reference resumes here and
one written here resumes in the reference — `itos.txt`, `curves.csv`, `meta.json`.

What is different: the step is `TrainStep` (flat fp32 buffers, fused AdamW, CUDA-graph replay for full micro-batches,
bucketed bf16 all-reduce), losses stay on the device and are read once per accumulation group, and with more than one
GPU `main()` spawns one process per GPU itself (rank 0 alone owns the run directory, guarded by a lock file).

Not rebuilt (out of scope): dataset preparation / manifests / vocabulary provenance, the DNA-shape encoder
pre-training, Adafactor, ReduceLROnPlateau, wall-time limits, MPS autocast.
"""
from __future__ import annotations

import argparse
import csv
import json
import math
import os
import time
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .model_tiny_gpt import TinyGPT
from .objectives import training_loss
from .token_feed import DeviceTokenStore, bucket_batches, rank_microbatches
from .trainer import (AccumulationHealth, NonfiniteGroupLimitError, TrainStep, cosine_lr_scale, resolve_warmup_steps,
                      run_accumulation_groups)

PAD_ID = 0


# ----------------------------------------------------------------------------------------------------------
# configuration -> model / step (same keys, same defaults as the reference)
# ----------------------------------------------------------------------------------------------------------
def normalize_offset_weights(targets: Sequence[int], weights) -> Dict[int, float]:
    """`multi_offset_weights` as the reference resolves it (training/config.py `_normalize_offset_weights`): a mapping
    offset -> weight, a list aligned with `multi_offset_targets`, or nothing (equal weights 1/len)."""
    targets = [int(t) for t in targets]
    if not targets:
        return {}
    if weights is None:
        return {t: 1.0 / len(targets) for t in targets}
    if isinstance(weights, dict):
        return {int(k): float(v) for k, v in weights.items() if int(k) in targets}
    weights = list(weights)
    if len(weights) != len(targets):
        raise ValueError("multi_offset_weights must match multi_offset_targets")
    return {t: float(w) for t, w in zip(targets, weights)}


def build_model_from_training_cfg(cfg: dict) -> TinyGPT:
    """loop.py:559-579."""
    multi = bool(cfg.get("multi_offset_loss_enabled", False))
    term = bool(cfg.get("termination_loss_enabled", False)) or bool(cfg.get("replay_loss_enabled", False))
    edges = tuple(int(x) for x in cfg.get("termination_bucket_edges", [0, 3, 10, 30]))
    return TinyGPT(
        cfg["vocab_size"], cfg["block_size"], n_layer=cfg["n_layer"], n_head=cfg["n_head"], n_embd=cfg["n_embd"],
        dropout=cfg["dropout"], use_checkpoint=bool(cfg.get("use_checkpoint", cfg.get("grad_checkpointing", False))),
        label_smoothing=float(cfg.get("label_smoothing", 0.0)),
        sep_id=(3 if bool(cfg.get("sep_mask_enabled", True)) else None),
        tie_embeddings=bool(cfg.get("tie_embeddings", True)),
        n_kv_head=int(cfg.get("n_kv_head")) if cfg.get("n_kv_head") is not None else None,
        use_sdpa=bool(cfg.get("use_sdpa", False)), loss_weights=cfg.get("loss_weights"),
        termination_aux=term, termination_n_classes=int(cfg.get("termination_n_classes", len(edges) + 1)),
        multi_offset_targets=[int(x) for x in cfg.get("multi_offset_targets", [])] if multi else None,
        use_swiglu=bool(cfg.get("use_swiglu", False)), use_rope=bool(cfg.get("use_rope", False)),
        use_shape_guidance=bool(cfg.get("use_shape_guidance", False)))


# ----------------------------------------------------------------------------------------------------------
# data: the reference's static NPZ format (X, Y int arrays [N, T]) resident in HBM, or the dynamic packed format
# ----------------------------------------------------------------------------------------------------------
class StaticTokenSet:
    """PackedDataset's static format (data_loading.py: NPZ with `X`, `Y`): both arrays uploaded once."""

    def __init__(self, X: np.ndarray, Y: np.ndarray, device):
        self.X = torch.from_numpy(np.ascontiguousarray(X).astype(np.int64)).to(device)
        self.Y = torch.from_numpy(np.ascontiguousarray(Y).astype(np.int64)).to(device)
        self.n = int(self.X.shape[0])

    def __len__(self):
        return self.n

    def batch(self, indices):
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64), device=self.X.device)
        return self.X.index_select(0, idx), self.Y.index_select(0, idx)

    @classmethod
    def from_npz(cls, paths: Sequence[str], device):
        xs, ys = [], []
        for p in paths:
            z = np.load(p)
            xs.append(z["X"])
            ys.append(z["Y"])
        return cls(np.concatenate(xs), np.concatenate(ys), device)


def epoch_order(n: int, batch_size: int, seed: Optional[int], shuffle: bool = True) -> List[List[int]]:
    """Index batches of one epoch exactly as `DataLoader(dataset, batch_size, shuffle=True, generator=g)` with
    `g.manual_seed(seed)` yields them (data_loading.py:465-478): the loader iterator first draws its base seed from
    the generator, then RandomSampler draws `randperm(n)` from the same generator; the last batch may be short."""
    if shuffle:
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(int(seed))
        torch.empty((), dtype=torch.int64).random_(generator=g)  # _BaseDataLoaderIter's base seed
        perm = torch.randperm(n, generator=g).tolist()
    else:
        perm = list(range(n))
    return [perm[i:i + batch_size] for i in range(0, n, batch_size)]


# ----------------------------------------------------------------------------------------------------------
# run directory (rank 0 only)
# ----------------------------------------------------------------------------------------------------------
class RunDir:
    """runs/<RUN_ID>/ as the reference lays it out: checkpoints/{last,best,best_epoch_XXX}.pt + config.yaml + meta.json,
    scores/curves.csv, itos.txt; one writer at a time (`.run.lock` held with flock for the lifetime of the run, as
    src/training/run_lifecycle.py:250-261 does)."""

    def __init__(self, root: Path, run_id: str):
        import fcntl
        self.path = Path(root) / run_id
        self.ckpt = self.path / "checkpoints"
        self.scores = self.path / "scores"
        self.ckpt.mkdir(parents=True, exist_ok=True)
        self.scores.mkdir(parents=True, exist_ok=True)
        self._lock = open(self.path / ".run.lock", "w")
        try:
            fcntl.flock(self._lock, fcntl.LOCK_EX | fcntl.LOCK_NB)
        except OSError as exc:
            raise RuntimeError(f"run directory {self.path} is locked by another training process") from exc

    def close(self):
        import fcntl
        try:
            fcntl.flock(self._lock, fcntl.LOCK_UN)
        finally:
            self._lock.close()


def save_checkpoint_atomic(payload: dict, path: Path):
    tmp = Path(str(path) + ".tmp")
    torch.save(payload, tmp)
    os.replace(tmp, path)


def capture_rng_state() -> dict:
    state = {"torch": torch.get_rng_state(), "numpy": np.random.get_state()}
    if torch.cuda.is_available():
        state["cuda"] = torch.cuda.get_rng_state_all()
    return state


# ----------------------------------------------------------------------------------------------------------
# the loop
# ----------------------------------------------------------------------------------------------------------
class Trainer:
    def __init__(self, cfg: dict, train_set, val_set, run_root: Optional[str] = None, run_id: Optional[str] = None,
                 device=None, process_group=None, log=print):
        self.cfg = cfg
        self.log = log
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.world = dist.get_world_size(process_group) if (process_group is not None or (
            dist.is_available() and dist.is_initialized())) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.pg = process_group
        self.train_set, self.val_set = train_set, val_set
        self.base_seed = int(cfg.get("seed", 1337))
        torch.manual_seed(self.base_seed)  # loop.py:289-291: the constructor's RNG contract gives the reference's init
        self.model = build_model_from_training_cfg(cfg).to(self.device)
        if bool(cfg.get("freeze_backbone", False)):  # loop.py:656-667
            for name, p in self.model.named_parameters():
                p.requires_grad = ("offset_projs" in name) or ("termination_head" in name)
        multi = bool(cfg.get("multi_offset_loss_enabled", False))
        self.offset_weights = normalize_offset_weights(cfg.get("multi_offset_targets", []),
                                                       cfg.get("multi_offset_weights")) if multi else {}
        self.term_enabled = bool(cfg.get("termination_loss_enabled", False))
        self.term_weight = float(cfg.get("termination_loss_weight", 0.1)) if self.term_enabled else 0.0
        self.term_stop_ids = tuple(int(x) for x in cfg.get("termination_stop_ids", [2]))
        self.term_edges = tuple(int(x) for x in cfg.get("termination_bucket_edges", [0, 3, 10, 30]))
        tcw = cfg.get("termination_class_weights")
        self.term_class_weights = None if tcw is None else torch.tensor([float(v) for v in tcw], device=self.device)
        lr = float(cfg.get("lr", 5e-6))
        self.step_obj = TrainStep(self.model, lr=lr, lr_embedding=float(cfg.get("lr_embedding", lr)),
                                  weight_decay=float(cfg.get("weight_decay", 0.05)),
                                  offset_weights=self.offset_weights or None,
                                  termination_loss_weight=self.term_weight, process_group=process_group,
                                  termination_stop_ids=self.term_stop_ids, termination_bucket_edges=self.term_edges,
                                  termination_class_weights=self.term_class_weights)
        self.gacc = int(cfg.get("grad_accum_steps", 16))
        if self.gacc % self.world != 0:
            raise ValueError("grad_accum_steps must be a multiple of the number of GPUs")
        self.batch_size = int(cfg["batch_size"])
        self.max_epochs = int(cfg.get("epochs", 5))
        n_batches = math.ceil(len(train_set) / self.batch_size)
        self.steps_per_epoch = math.ceil(n_batches / max(1, self.gacc))
        self.total_steps = int(cfg.get("scheduler_total_steps", max(1, self.steps_per_epoch * self.max_epochs)))
        self.warmup_steps = resolve_warmup_steps(cfg, self.total_steps)
        cfg["resolved_warmup_steps"] = self.warmup_steps
        base_lr = float(cfg["lr"])
        self.min_lr_ratio = (float(cfg.get("min_lr", 1e-5)) / base_lr) if base_lr > 0 else 0.0
        self.health = AccumulationHealth()
        self.max_nonfinite = int(cfg.get("max_nonfinite_accumulation_groups", 3))
        self.step = 0
        self.start_epoch = 0
        self.best, self.best_epoch, self.no_improve = float("inf"), None, 0
        self.consumed_train_tokens = 0
        self.history: List[dict] = []
        self.run: Optional[RunDir] = None
        if self.rank == 0 and run_root is not None:
            self.run = RunDir(Path(run_root), run_id or time.strftime("%Y-%m-%d_%H%M%S"))
            itos = cfg.get("itos_path")
            if itos and Path(itos).exists():
                (self.run.path / "itos.txt").write_text(Path(itos).read_text())
            self.log_csv = self.run.scores / "curves.csv"
            if not self.log_csv.exists():
                with self.log_csv.open("w", newline="") as f:
                    csv.writer(f).writerow(["step", "train_loss", "val_loss", "train_next_loss", "val_next_loss",
                                            "perplexity", "lr"])
            try:
                import yaml
                (self.run.ckpt / "config.yaml").write_text(yaml.safe_dump({k: v for k, v in cfg.items()
                                                                          if isinstance(v, (int, float, str, bool, list, dict, type(None)))}))
            except Exception:  # pragma: no cover - the YAML copy is a convenience for the reference's tools
                pass

    # ------------------------------------------------------------------ schedule
    def lr_scale(self, step_idx: int) -> float:
        return cosine_lr_scale(step_idx, self.warmup_steps, self.total_steps, self.min_lr_ratio)

    # ------------------------------------------------------------------ data
    def _epoch_batches(self, epoch_idx: int) -> List[List[int]]:
        seed = self.base_seed + max(0, int(epoch_idx))  # loop.py `_loader_cfg_for_epoch(epoch + 1)`
        if isinstance(self.train_set, DeviceTokenStore) and bool(self.cfg.get("bucket_batching", False)):
            return bucket_batches(self.train_set.lengths_host, self.batch_size, int(self.cfg.get("n_buckets", 8)),
                                  shuffle=True, seed=seed)
        return epoch_order(len(self.train_set), self.batch_size, seed)

    def _fetch(self, data, indices):
        if isinstance(data, DeviceTokenStore):
            return data.fetch_batch(np.asarray(indices, dtype=np.int64))
        return data.batch(indices)

    # ------------------------------------------------------------------ passes
    def _loss(self, xb, yb):
        return training_loss(self.model, xb, yb, offset_weights=self.offset_weights or None,
                             termination_loss_weight=self.term_weight, termination_stop_ids=self.term_stop_ids,
                             termination_bucket_edges=self.term_edges,
                             termination_class_weights=self.term_class_weights)

    def train_epoch(self, epoch_idx: int, skip_microbatches: int = 0, on_microbatch=None) -> Tuple[float, float]:
        self.model.train()
        batches = self._epoch_batches(epoch_idx)
        if skip_microbatches:
            batches = batches[skip_microbatches:]
        mine = rank_microbatches(batches, self.rank, self.world, self.gacc) if self.world > 1 else iter(batches)

        def stream() -> Iterator:
            for ind in mine:
                xb, yb = self._fetch(self.train_set, ind)
                if on_microbatch is not None:
                    on_microbatch(ind)
                yield xb, yb
        total_sum = next_sum = 0.0
        n = 0
        for rec in run_accumulation_groups(self.step_obj, stream(), self.gacc // self.world, self.health,
                                           max_nonfinite_groups=self.max_nonfinite, lr_scale_fn=self.lr_scale,
                                           first_step_idx=self.step, process_group=self.pg):
            self.step = rec["step"] + 1
            total_sum += rec["total_loss_sum"]
            next_sum += rec["next_loss_sum"]
            n += rec["group_size"]
        if self.world > 1:  # the logged epoch means cover every rank's micro-batches
            t = torch.tensor([total_sum, next_sum, float(n)], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            total_sum, next_sum, n = float(t[0]), float(t[1]), int(round(float(t[2])))
        return total_sum / max(n, 1), next_sum / max(n, 1)

    @torch.no_grad()
    def validate(self) -> Tuple[float, float]:
        self.model.eval()
        totals, nexts = [], []
        for ind in epoch_order(len(self.val_set), self.batch_size, None, shuffle=False):
            xb, yb = self._fetch(self.val_set, ind)
            total, parts, _ = self._loss(xb, yb)
            totals.append(total.reshape(1))
            nexts.append(parts["next"].reshape(1))
        t = torch.cat(totals + nexts).float().cpu()  # one host read for the whole pass
        k = len(totals)
        finite = torch.isfinite(t[:k])
        if not bool(finite.any()):
            return float("nan"), float("nan")
        return float(t[:k][finite].mean()), float(t[k:][finite].mean())

    # ------------------------------------------------------------------ checkpoints (loop.py:950-1007)
    def checkpoint_payload(self, epoch_idx: int, train_loss=float("inf"), val_loss=float("inf"), train_next_loss=None,
                           val_next_loss=None, microbatch_idx: int = 0) -> dict:
        n_batches = math.ceil(len(self.train_set) / self.batch_size)
        return {
            "model": {k: v.detach().clone() for k, v in self.model.state_dict().items()},
            "optimizer": self.step_obj.state_dict(lr_scale_fn=self.lr_scale),
            "scheduler": self.step_obj.scheduler_state_dict(self.lr_scale),
            "cfg": self.cfg,
            "epoch": max(0, epoch_idx - 1) if val_loss == float("inf") else epoch_idx,
            "val_loss": val_loss, "train_loss": train_loss, "train_next_loss": train_next_loss,
            "val_next_loss": val_next_loss, "train_term_loss": None, "val_term_loss": None,
            "train_replay_term_loss": None,
            "best_val": self.best, "best_epoch": self.best_epoch, "no_improve": self.no_improve, "step": self.step,
            "consumed_train_tokens": int(self.consumed_train_tokens), "runtime_memory": {},
            "epoch_microbatch_idx": 0 if val_loss != float("inf") else int(microbatch_idx),
            "last_seen_microbatch_idx": int(microbatch_idx),
            "batch_size": self.batch_size, "grad_accum_steps": self.gacc, "train_examples": int(len(self.train_set)),
            "train_batches": int(n_batches), "accumulation_health": self.health.state_dict(),
            "max_nonfinite_accumulation_groups": self.max_nonfinite, "epoch_train_metrics": {},
            "run_progress": {"completed_epochs": epoch_idx if val_loss != float("inf") else max(0, epoch_idx - 1),
                             "current_epoch": epoch_idx, "microbatch": 0, "optimizer_step": self.step},
            "rng_state": capture_rng_state(), "run_fingerprint": None,
        }

    def resume(self, path: str):
        """loop.py:880-940: model (strict), optimiser, scheduler position, counters — from a `last.pt` written here or
        by the reference trainer."""
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ck["model"])
        if "optimizer" in ck:
            self.step_obj.load_state_dict(ck["optimizer"])
        self.start_epoch = int(ck.get("epoch", 0))
        self.step = int(ck.get("step", 0))
        self.step_obj.step_count = self.step if ck.get("optimizer", {}).get("state") else self.step_obj.step_count
        self.consumed_train_tokens = int(ck.get("consumed_train_tokens", 0))
        self.best = float(ck.get("best_val", self.best))
        self.best_epoch = ck.get("best_epoch", self.best_epoch)
        self.no_improve = int(ck.get("no_improve", 0))
        self.health.load_state_dict(ck.get("accumulation_health"))
        self.log(f"[resume] loaded {path}: epoch {self.start_epoch}, optimizer step {self.step}")
        return ck

    def _save(self, payload, names: Sequence[str]):
        if self.run is None:
            return
        for name in names:
            save_checkpoint_atomic(payload, self.run.ckpt / name)

    # ------------------------------------------------------------------ fit
    def fit(self) -> dict:
        patience = int(self.cfg.get("early_stop_patience", 5))
        t0 = time.perf_counter()
        status = "completed"
        try:
            for epoch in range(self.start_epoch, self.max_epochs):
                epoch_idx = epoch + 1
                train_loss, train_next = self.train_epoch(epoch_idx)
                val_loss, val_next = self.validate()
                ppl = math.exp(min(20.0, val_next)) if math.isfinite(val_next) else float("nan")
                lr_now = self.step_obj.groups[0].lr * self.lr_scale(self.step)
                self.log(f"[epoch {epoch_idx}] train {train_loss:.3f} | val {val_loss:.3f} | next_val {val_next:.3f} "
                         f"| ppl {ppl:.2f} | lr {lr_now:.2e}")
                improved = val_loss + 1e-6 < self.best
                if improved:
                    self.best, self.best_epoch, self.no_improve = val_loss, epoch_idx, 0
                else:
                    self.no_improve += 1
                self.history.append({"epoch": epoch_idx, "train_loss": train_loss, "val_loss": val_loss,
                                     "train_next_loss": train_next, "val_next_loss": val_next, "perplexity": ppl,
                                     "lr": lr_now, "step": self.step, **self.health.metrics_dict()})
                if self.rank == 0 and self.run is not None:
                    payload = self.checkpoint_payload(epoch_idx, train_loss, val_loss, train_next, val_next)
                    names = ["last.pt"] + ([f"epoch_{epoch_idx}.pt"] if self.cfg.get("save_epochs", False) else [])
                    if improved:
                        names += ["best.pt", f"best_epoch_{epoch_idx:03d}.pt"]
                    self._save(payload, names)
                    with self.log_csv.open("a", newline="") as f:
                        csv.writer(f).writerow([epoch_idx, f"{train_loss:.4f}", f"{val_loss:.4f}", f"{train_next:.4f}",
                                                f"{val_next:.4f}", f"{ppl:.3f}", f"{lr_now:.3e}"])
                if not improved and patience > 0 and self.no_improve >= patience:
                    self.log("[early-stopping] no improvement; stopping.")
                    break
        except NonfiniteGroupLimitError:
            status = "failed"
            if self.rank == 0 and self.run is not None:
                payload = self.checkpoint_payload(self.start_epoch + 1)
                payload["checkpoint_reason"] = "nonfinite_group_limit"
                self._save(payload, ["last.pt"])
            raise
        finally:
            if self.rank == 0 and self.run is not None:
                meta = {"run_id": self.run.path.name, "train_wall_sec": round(time.perf_counter() - t0, 2),
                        "best_epoch": self.best_epoch,
                        "best_val_loss": float(self.best) if self.best != float("inf") else None, "status": status,
                        "accumulation_health": self.health.state_dict(), "model_spec": self.model.to_dict(),
                        "history": self.history}
                (self.run.ckpt / "meta.json").write_text(json.dumps(meta, indent=2, default=str))
                self.run.close()
        return {"history": self.history, "step": self.step, "best_val": self.best, "best_epoch": self.best_epoch}


# ----------------------------------------------------------------------------------------------------------
# CLI: the reference's `python -m src.codonlm.train_codon_lm` arguments; spawns one process per GPU itself
# ----------------------------------------------------------------------------------------------------------
def _worker(rank: int, world: int, port: int, cfg: dict, args):
    pg = None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        import datetime
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                                timeout=datetime.timedelta(seconds=int(cfg.get("collective_timeout_s", 600))))
        pg = dist.group.WORLD
    dev = torch.device("cuda", rank)
    train = StaticTokenSet.from_npz(args.train_npz, dev)
    val = StaticTokenSet.from_npz(args.val_npz, dev)
    run_id = args.run_id or os.environ.get("RUN_ID") or cfg.get("run_id")
    tr = Trainer(cfg, train, val, run_root=cfg.get("runs_dir", "runs"), run_id=run_id, device=dev, process_group=pg,
                 log=(print if rank == 0 else (lambda *a, **k: None)))
    if args.resume:
        tr.resume(args.resume)
    try:
        tr.fit()
    finally:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


def main(argv=None):
    import yaml
    ap = argparse.ArgumentParser(description="codon-LM training on B200 (arguments of src/codonlm/train_codon_lm.py)")
    ap.add_argument("--config", required=True)
    ap.add_argument("--run_id", default=None)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--train_npz", action="append", default=None)
    ap.add_argument("--val_npz", action="append", default=None)
    ap.add_argument("--test_npz", action="append", default=None)
    ap.add_argument("--save_epochs", action="store_true")
    ap.add_argument("--gpus", type=int, default=None, help="GPUs of this box to train on (default: all visible)")
    args = ap.parse_args(argv)
    cfg = yaml.safe_load(open(args.config)) or {}
    if isinstance(cfg.get("data"), dict):
        for k, v in cfg["data"].items():
            cfg.setdefault(k, v)
    cfg["save_epochs"] = args.save_epochs or cfg.get("save_epochs", False)
    args.train_npz = args.train_npz or cfg.get("train_npz")
    args.val_npz = args.val_npz or cfg.get("val_npz")
    if isinstance(args.train_npz, str):
        args.train_npz = [args.train_npz]
    if isinstance(args.val_npz, str):
        args.val_npz = [args.val_npz]
    if not args.train_npz or not args.val_npz:
        ap.error("--train_npz and --val_npz (or the config's train_npz / val_npz) are required")
    if not torch.cuda.is_available():
        raise SystemExit("codonlm_b200.train: no CUDA device (this trainer has no CPU path)")
    world = args.gpus or torch.cuda.device_count()
    gacc = int(cfg.get("grad_accum_steps", 16))
    while world > 1 and gacc % world != 0:
        world -= 1  # the accumulation group must deal evenly over the ranks
    if world <= 1:
        _worker(0, 1, 0, cfg, args)
        return
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, cfg, args), nprocs=world, join=True)


if __name__ == "__main__":
    main()
