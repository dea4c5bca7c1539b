"""Token feed: the oracle restatement of MmapPackedDataset.fetch_batch and the batch order of
BucketBatchSampler against golden vectors produced by the unmodified reference (tests/golden/make_feed_golden.py);
the device-resident store (cgpt_pack_lm_batch) against the same vectors, bit-exact (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import codon_gpt_oracle as O

GOLD = os.path.join(ROOT, "tests", "golden", "token_feed.npz")


def _load():
    z = np.load(GOLD)
    flats = [z["flat0"], z["flat1"]]
    lens = [z["lengths0"], z["lengths1"]]
    return z, flats, lens


def test_oracle_fetch_batch_matches_reference_golden():
    z, flats, lens = _load()
    flat, lengths = np.concatenate(flats), np.concatenate(lens)
    for i in range(int(z["n_batches"])):
        xb, yb = O.fetch_batch_dynamic(flat, lengths, z[f"idx{i}"])
        assert xb.dtype == np.int64 and np.array_equal(xb, z[f"xb{i}"]) and np.array_equal(yb, z[f"yb{i}"]), i
    xb, yb = O.fetch_batch_dynamic(flat, lengths, [])
    assert xb.shape == (0, 0) and yb.shape == (0, 0)


def test_bucket_batches_follow_reference_sampler_order():
    from codonlm_b200.token_feed import bucket_batches
    z, _, lens = _load()
    got = bucket_batches(np.concatenate(lens), batch_size=7, n_buckets=4, shuffle=True, drop_last=False, seed=11)
    assert [len(b) for b in got] == z["sampler_sizes"].tolist()
    assert np.concatenate([np.asarray(b) for b in got]).tolist() == z["sampler_flat"].tolist()


def test_rank_microbatches_partition_every_group():
    from codonlm_b200.token_feed import rank_microbatches
    batches = list(range(19))  # 4 full groups of 4 + a partial one
    parts = [list(rank_microbatches(batches, r, 2, 4)) for r in range(2)]
    assert parts[0] == [0, 2, 4, 6, 8, 10, 12, 14, 16, 18] and parts[1] == [1, 3, 5, 7, 9, 11, 13, 15, 17]
    assert sorted(parts[0] + parts[1]) == batches
    with pytest.raises(ValueError):
        list(rank_microbatches(batches, 0, 3, 4))


@pytest.mark.gpu
def test_device_token_store_is_bit_exact():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from codonlm_b200.token_feed import DeviceTokenStore
    z, flats, lens = _load()
    store = DeviceTokenStore(list(zip(flats, lens)))
    assert len(store) == sum(len(l) for l in lens)
    for i in range(int(z["n_batches"])):
        xb, yb = store.fetch_batch(z[f"idx{i}"])
        assert xb.dtype == torch.int64 and xb.is_cuda
        assert np.array_equal(xb.cpu().numpy(), z[f"xb{i}"]) and np.array_equal(yb.cpu().numpy(), z[f"yb{i}"]), i
    xb, yb = store.fetch_batch([])
    assert xb.shape == (0, 0)
    # bucketed widths: same content, PAD columns appended
    wide = DeviceTokenStore(list(zip(flats, lens)), bucket=32)
    xb, yb = wide.fetch_batch(z["idx2"])
    ref = z["xb2"]
    assert xb.shape[1] % 32 == 0 and xb.shape[1] >= ref.shape[1]
    assert np.array_equal(xb.cpu().numpy()[:, : ref.shape[1]], ref) and int(xb[:, ref.shape[1]:].abs().sum()) == 0
    assert np.array_equal(yb.cpu().numpy()[:, : ref.shape[1]], z["yb2"]) and int(yb[:, ref.shape[1]:].abs().sum()) == 0
    # a full-size check against the oracle: 4096 sequences of up to 1025 tokens
    rng = np.random.default_rng(0)
    lengths = rng.integers(2, 1026, size=4096).astype(np.int64)
    flat = rng.integers(4, 68, size=int(lengths.sum())).astype(np.int32)
    big = DeviceTokenStore([(flat, lengths)])
    idx = rng.choice(4096, size=64, replace=False)
    xb, yb = big.fetch_batch(idx)
    rx, ry = O.fetch_batch_dynamic(flat, lengths, idx)
    assert np.array_equal(xb.cpu().numpy(), rx) and np.array_equal(yb.cpu().numpy(), ry)
