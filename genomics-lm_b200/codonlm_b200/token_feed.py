"""Token feed for the training step: the reference's packed dataset held RESIDENT in HBM.

The reference gathers every micro-batch on the host (`MmapPackedDataset.fetch_batch`, src/codonlm/data_loading.py
:271-315: numpy row gather, shift, pad to the longest sequence of the batch) and copies (xb, yb) int64 to the device.
Its datasets are a few MB to a few GB of small integers, a B200 has 180 GB: here the flat token array is uploaded once
(int32) with its offsets / lengths, and a batch is built by ONE kernel from the B sequence indices
(`cgpt_pack_lm_batch`) — per step only 8·B bytes cross PCIe.  Output is bit-identical to `fetch_batch`.

`bucket` (optional) rounds the batch width up to a multiple, so that a handful of (B, T) shapes cover an epoch and
the CUDA-graph step (`TrainStep.capture`) can be reused; the extra columns are PAD and change neither the loss
(ignore_index 0) nor any visible attention.  `rank_microbatches` deals micro-batches r, r+W, ... of every
accumulation group to rank r, keeping the reference's global batch order.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np
import torch

from . import ops


class DeviceTokenStore:
    """Dynamic-format shards (flat `X`, `lengths`) concatenated and resident on the device."""

    def __init__(self, shards: Sequence[Tuple[np.ndarray, np.ndarray]], device="cuda", bucket: int = 1):
        flats, lens = [], []
        for flat, lengths in shards:
            flat, lengths = np.asarray(flat), np.asarray(lengths, dtype=np.int64)
            if int(lengths.sum()) != flat.shape[0]:
                raise ValueError("lengths do not add up to the size of the flat token array")
            flats.append(flat.astype(np.int32, copy=False))
            lens.append(lengths)
        self.lengths_host = np.concatenate(lens) if lens else np.zeros((0,), np.int64)
        offsets = np.concatenate([[0], np.cumsum(self.lengths_host[:-1])]).astype(np.int64) if len(self.lengths_host) \
            else np.zeros((0,), np.int64)
        self.device = torch.device(device)
        self.tokens = torch.from_numpy(np.concatenate(flats) if flats else np.zeros((0,), np.int32)).to(self.device)
        self.offsets = torch.from_numpy(offsets).to(self.device)
        self.lengths = torch.from_numpy(self.lengths_host).to(self.device)
        self.bucket = max(1, int(bucket))

    def __len__(self) -> int:
        return int(self.lengths_host.shape[0])

    @property
    def seq_lengths(self) -> np.ndarray:  # what BucketBatchSampler reads (data_loading.py:317-329)
        return self.lengths_host

    def batch_width(self, indices: np.ndarray) -> int:
        target = max(0, int(self.lengths_host[indices].max()) - 1)
        return (target + self.bucket - 1) // self.bucket * self.bucket

    def fetch_batch(self, indices) -> Tuple[torch.Tensor, torch.Tensor]:
        """(xb, yb) int64 on the device, equal to MmapPackedDataset.fetch_batch(indices) (padded to `bucket`)."""
        indices = np.asarray(indices, dtype=np.int64)
        if indices.size == 0:
            empty = torch.empty((0, 0), dtype=torch.long, device=self.device)
            return empty, empty.clone()
        width = self.batch_width(indices)
        if width == 0:
            empty = torch.empty((indices.size, 0), dtype=torch.long, device=self.device)
            return empty, empty.clone()
        idx_dev = torch.from_numpy(indices).to(self.device, non_blocking=True)
        return ops.pack_lm_batch(self.tokens, self.offsets, self.lengths, idx_dev, width)


def rank_microbatches(batches: Iterable, rank: int, world: int, grad_accum_steps: int) -> Iterator:
    """Micro-batches r, r+W, ... of every accumulation group of `grad_accum_steps` consecutive batches go to rank r,
    so that one optimiser step consumes the same sequences as the single-process reference.
    `grad_accum_steps` must be a multiple of `world`."""
    if grad_accum_steps % world != 0:
        raise ValueError("grad_accum_steps must be a multiple of the world size")
    for i, b in enumerate(batches):
        if (i % grad_accum_steps) % world == rank:
            yield b


def bucket_batches(lengths: np.ndarray, batch_size: int, n_buckets: int = 8, shuffle: bool = True,
                   drop_last: bool = False, seed=None) -> List[List[int]]:
    """Length-bucketed batches in the order of the reference's BucketBatchSampler (data_loading.py:332-368): same
    bucket edges, same generator, same two shuffles — one epoch's list of index batches."""
    lengths = np.asarray(lengths)
    edges = np.linspace(lengths.min(), lengths.max() + 1, n_buckets + 1)
    bucket_ids = np.digitize(lengths, edges[1:])
    buckets: List[List[int]] = [[] for _ in range(n_buckets)]
    for idx, bid in enumerate(bucket_ids):
        buckets[bid].append(idx)
    rng = np.random.default_rng(seed)
    out: List[List[int]] = []
    for bucket in buckets:
        if not bucket:
            continue
        order = list(bucket)
        if shuffle:
            rng.shuffle(order)
        for start in range(0, len(order), batch_size):
            chunk = order[start:start + batch_size]
            if drop_last and len(chunk) < batch_size:
                continue
            out.append(chunk)
    if shuffle:
        rng.shuffle(out)
    return out
