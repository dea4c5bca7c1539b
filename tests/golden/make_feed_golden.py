"""Golden vectors for the token feed, produced by the UNMODIFIED reference (run in the build container):
  PYTHONPATH=/root/reference python tests/golden/make_feed_golden.py
Writes tests/golden/token_feed.npz: a small dynamic-format dataset (two shards: flat X + lengths), index batches
(including a batch holding a length-1 sequence) with the (xb, yb) that MmapPackedDataset.fetch_batch returns
(src/codonlm/data_loading.py:271-315), and one epoch of BucketBatchSampler batches (:332-368)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, "/root/reference")
from src.codonlm.data_loading import BucketBatchSampler, MmapPackedDataset  # noqa: E402

rng = np.random.default_rng(20251018)
out = {}
paths = []
tmp = tempfile.mkdtemp()
for shard in range(2):
    lengths = rng.integers(1, 90, size=40 + 10 * shard).astype(np.int64)
    lengths[3] = 1  # a sequence with no usable position
    flat = rng.integers(4, 68, size=int(lengths.sum())).astype(np.int16)
    path = os.path.join(tmp, f"shard{shard}.npz")
    np.savez(path, X=flat, lengths=lengths)
    paths.append(path)
    out[f"flat{shard}"] = flat
    out[f"lengths{shard}"] = lengths
ds = MmapPackedDataset(paths)
assert ds.is_dynamic and ds.supports_batched_fetch
batches = [rng.choice(len(ds), size=n, replace=False) for n in (1, 5, 16, 33)] + [np.array([3]), np.array([3, 43])]
for i, idx in enumerate(batches):
    xb, yb = ds.fetch_batch(idx)
    out[f"idx{i}"] = np.asarray(idx, dtype=np.int64)
    out[f"xb{i}"] = xb.numpy()
    out[f"yb{i}"] = yb.numpy()
out["n_batches"] = np.array(len(batches))
sampler = BucketBatchSampler(ds.seq_lengths, batch_size=7, n_buckets=4, shuffle=True, drop_last=False, seed=11)
epoch = list(sampler)
out["sampler_sizes"] = np.array([len(b) for b in epoch])
out["sampler_flat"] = np.concatenate([np.asarray(b, dtype=np.int64) for b in epoch])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "token_feed.npz"), **out)
print("wrote token_feed.npz:", len(ds), "sequences,", len(batches), "batches,", len(epoch), "sampler batches")
