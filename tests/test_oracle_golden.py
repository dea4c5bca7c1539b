"""Pin the CPU oracle against the reference's own outputs (tests/golden, made by make_golden.py)
and against the reference's known-answer tables."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, load_golden
from oracle import codon_gpt_oracle as O


def _cfg(meta):
    return O.make_cfg(**meta["ctor"])


def test_forward_matches_reference(golden):
    case, z, meta, sd, grads = golden
    cfg = _cfg(meta)
    idx, tgt = torch.from_numpy(z["idx"]), torch.from_numpy(z["targets"])
    shapes = torch.from_numpy(z["shape_embeddings"]) if "shape_embeddings" in z.files else None
    out = O.forward(sd, cfg, idx, tgt, attention_window=meta["attention_window"], want_hidden=True,
                    shape_embeddings=shapes)
    # fp32 on both sides, different op order (manual softmax vs SDPA): 1e-4 abs at |logit| <~ 50
    scale = max(1.0, float(np.abs(z["logits"]).max()))
    assert np.abs(out["logits"].numpy() - z["logits"]).max() <= 2e-5 * scale
    assert out["loss"].item() == pytest.approx(meta["parts"]["next"], rel=2e-6)
    assert np.array_equal(out["logits"].argmax(-1).numpy(), z["argmax"])
    # the embedding gather is exact; with shape guidance the K=3 projection is added (different summation order)
    assert np.abs(out["hidden"][0].numpy() - z["hidden_0"]).max() <= (0.0 if shapes is None else 1e-6)
    assert np.abs(out["hidden"][-1].numpy() - z["hidden_final"]).max() <= 2e-5
    if "termination_logits" in z.files:
        assert np.abs(out["termination_logits"].numpy() - z["termination_logits"]).max() <= 2e-5
    for o, lg in out.get("offset_logits", {}).items():
        assert np.abs(lg.numpy() - z[f"offset_logits.{o}"]).max() <= 2e-5


def test_mask_matches_reference(golden):
    case, z, meta, sd, grads = golden
    m = O.attention_mask(z["idx"], meta["ctor"].get("sep_id", 3), meta["attention_window"])
    if m is None:
        assert z["attn_mask"].size == 0
    else:
        assert np.array_equal(m, z["attn_mask"])


def test_losses_and_grads_match_reference(golden):
    case, z, meta, sd, grads = golden
    cfg = _cfg(meta)
    idx, tgt = torch.from_numpy(z["idx"]), torch.from_numpy(z["targets"])
    ow = {int(k): v for k, v in meta["offset_weights"].items()} or None
    shapes = None
    if "shape_embeddings" in z.files:
        shapes = torch.from_numpy(z["shape_embeddings"]).requires_grad_(True)
    total, parts, out, g = O.loss_and_grads(sd, cfg, idx, tgt, offset_weights=ow,
                                            termination_loss_weight=meta["termination_loss_weight"],
                                            attention_window=meta["attention_window"], shape_embeddings=shapes)
    if shapes is not None:
        ref = z["grad_shape_embeddings"]
        assert np.abs(shapes.grad.numpy() - ref).max() <= 2e-5 * np.abs(ref).max()
    assert total.item() == pytest.approx(meta["parts"]["total"], rel=3e-6)
    for o, v in meta["parts"].get("offsets", {}).items():
        assert parts["offsets"][int(o)].item() == pytest.approx(v, rel=3e-6)
    if "termination" in meta["parts"]:
        assert parts["termination"].item() == pytest.approx(meta["parts"]["termination"], rel=3e-6)
    assert set(g) == set(grads)
    gmax = max(v.norm().item() for v in grads.values())
    for k, ref in grads.items():
        num = (g[k] - ref).norm().item()
        den = ref.norm().item()
        # key.bias gradients are analytically 0 (softmax shift invariance): floor on the global scale
        assert num <= 2e-5 * den + 1e-6 * gmax, (k, num, den)


def test_integer_kats_from_reference_run():
    z = np.load(f"{GOLDEN_DIR}/integer_kats.npz")
    yb = z["yb"]
    for o in (1, 2, 3, 4, 8, 16, 32):
        assert np.array_equal(O.offset_target_mask(yb, o), z[f"offset_mask.{o}"])
    for name, (stops, edges) in {"a": ((2,), (0, 3, 10, 30)), "b": ((2, 3), (0, 1, 3)), "c": ((2,), ())}.items():
        assert np.array_equal(O.termination_distance_bucket_labels(yb, stops, edges), z[f"term.{name}"])


# ---- the reference's own known-answer tables, restated -------------------------------------

def test_mask_truth_table():  # reference tests/test_models.py:29-51
    tokens = np.array([[1, 4, 3, 5, 6]])
    full = O.attention_mask(tokens, 3)[0, 0]
    assert full[1, 0] and not full[3, 1] and full[3, 2] and full[4, 2]
    local = O.attention_mask(tokens, 3, attention_window=1)[0, 0]
    assert np.array_equal(local, np.eye(5, dtype=bool))
    with pytest.raises(ValueError, match="at least 1"):
        O.attention_mask(tokens, 3, attention_window=0)
    assert O.attention_mask(tokens, None) is None


def test_offset_mask_table():  # reference tests/test_long_range_codon_objectives.py:15-32
    yb = np.array([[10, 11, 12, 13, 0], [10, 2, 12, 13, 0], [10, 11, 3, 13, 0]])
    assert O.offset_target_mask(yb, 4, (2, 3)).tolist() == [[True, False], [False, False], [False, False]]
    assert O.offset_target_mask(yb, 6).shape == (3, 0)


def test_termination_label_table():  # reference tests/test_long_range_codon_objectives.py:72-92
    yb = np.array([[10, 11, 2, 12, 0], [2, 10, 11, 12, 13], [10, 11, 12, 13, 14]])
    lab = O.termination_distance_bucket_labels(yb, (2,), (0, 1, 3))
    assert lab.tolist() == [[2, 1, 0, 3, -100], [0, 3, 3, 3, 3], [3, 3, 3, 3, 3]]


def test_multi_offset_skips_when_no_valid_target():  # :35-42
    logits = {4: torch.randn(2, 4, 16)}
    total, losses = O.multi_offset_lm_loss(logits, torch.zeros((2, 4), dtype=torch.long), {4: 0.1})
    assert total.item() == 0.0 and losses == {}


def test_cross_entropy_equals_torch():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(50, 68, generator=g) * 3
    t = torch.randint(0, 68, (50,), generator=g)
    w = torch.rand(68, generator=g) + 0.5
    for eps in (0.0, 0.05):
        for ww in (None, w):
            a = O.cross_entropy(logits, t, 0, eps, ww)
            b = torch.nn.functional.cross_entropy(logits, t, ignore_index=0, label_smoothing=eps, weight=ww)
            assert a.item() == pytest.approx(b.item(), rel=1e-6)
    assert torch.isnan(O.cross_entropy(logits, torch.zeros(50, dtype=torch.long)))


def test_causality_and_segment_isolation():  # reference tests/test_embedding_extraction_contract.py:27-44
    cfg = O.make_cfg(68, 16, n_layer=2, n_head=2, n_embd=32, dropout=0.0)
    sd = O.init_state_dict(cfg, seed=3)
    a = torch.tensor([[1, 5, 6, 7, 3, 9, 10, 11]])
    b = a.clone()
    b[0, 6:] = torch.tensor([20, 21])
    ha = O.forward(sd, cfg, a, want_hidden=True)["hidden"][-1]
    hb = O.forward(sd, cfg, b, want_hidden=True)["hidden"][-1]
    assert torch.equal(ha[:, :6], hb[:, :6])
    c = a.clone()
    c[0, 1:4] = torch.tensor([30, 31, 32])  # other side of the <SEP> at position 4
    hc = O.forward(sd, cfg, c, want_hidden=True)["hidden"][-1]
    # tokens after the separator only differ through... nothing: pos-emb and own segment are unchanged
    assert torch.allclose(ha[:, 5:], hc[:, 5:], atol=0, rtol=0)


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
def test_oracle_at_baseline_shapes_matches_reference(name):
    """BASELINE.json configs[0] ('tiny 2L4H d128, seq 256, fp32 forward+loss on CPU'), configs[1] (6L4H d256
    RoPE+SwiGLU, seq 512), configs[2] (12L8H d512 with the multi-offset and termination heads, seq 1024) and configs[3]
    (bench_b8_gqa4) at full shape: the oracle rebuilds the weights the unmodified reference was run on
    (tests/golden/make_baseline_shape_golden.py) and must reproduce its loss, sampled logits, argmax map,
    hidden-state row norms and — for C3 — the offset and termination losses of the reference's objectives."""
    import json
    import os
    from conftest import ROOT
    z = np.load(os.path.join(ROOT, "tests", "golden", "baseline_shapes.npz"))
    meta = json.loads(str(z[name + ".meta"]))
    cfg = O.make_cfg(**meta["ctor"])
    sd = O.init_state_dict(cfg, seed=1337, emb_scale=0.02)
    idx, tgt = O.synthetic_batch(meta["B"], meta["T"], seed=1337, realistic=True)
    out = O.forward(sd, cfg, idx, tgt, want_hidden=True)
    assert out["loss"].item() == pytest.approx(float(z[name + ".loss"]), rel=3e-6)
    flat = out["logits"].reshape(-1)
    assert np.abs(flat[::97].numpy() - z[name + ".logit_samples"]).max() <= 3e-5
    assert float(flat.abs().mean()) == pytest.approx(float(z[name + ".logits_abs_mean"]), rel=1e-5)
    assert np.array_equal(out["logits"].argmax(-1).numpy().astype(np.int8), z[name + ".argmax"])
    norms = out["hidden"][-1].norm(dim=-1).numpy()
    assert np.abs(norms - z[name + ".hidden_row_norms"]).max() <= 1e-4 * np.abs(norms).max()
    # backward: per-parameter gradient norms and sampled entries of the trainer's total loss
    names = json.loads(str(z[name + ".grad_names"]))
    offs_all = meta["ctor"].get("multi_offset_targets")
    kw = dict(offset_weights={o: 0.2 for o in offs_all}, termination_loss_weight=0.1) if offs_all else {}
    total, _, _, grads = O.loss_and_grads(sd, cfg, idx, tgt, **kw)
    assert total.item() == pytest.approx(float(z[name + ".total_loss"]), rel=3e-6)
    assert set(names) == set(grads)
    ref_norms = z[name + ".grad_norms"]
    got_norms = np.array([grads[k].norm().item() for k in names])
    assert np.abs(got_norms - ref_norms).max() <= 3e-5 * ref_norms.max()
    samples = torch.cat([grads[k].reshape(-1)[::997] for k in names]).numpy()
    assert np.abs(samples - z[name + ".grad_samples"]).max() <= 3e-5 * np.abs(z[name + ".grad_samples"]).max()
    if name + ".offset_losses" in z.files:
        offs = meta["ctor"]["multi_offset_targets"]
        total, parts, _ = O.training_loss(sd, cfg, idx, tgt, offset_weights={o: 0.2 for o in offs},
                                          termination_loss_weight=0.1)
        got = np.array([parts["offsets"][o].item() for o in offs])
        assert np.abs(got - z[name + ".offset_losses"]).max() <= 3e-6 * np.abs(got).max()
        assert parts["termination"].item() == pytest.approx(float(z[name + ".termination_loss"]), rel=3e-6)
        o32 = out["offset_logits"][32].reshape(-1)[::97].numpy()
        assert np.abs(o32 - z[name + ".offset32_logit_samples"]).max() <= 3e-5


def test_replay_term_of_the_trainer_loss_matches_reference():
    """The replay branch of fwd() (loop.py:1113-1141): second forward over generated contexts, termination-head CE on
    sparse labels with replay_class_weights, added with replay_loss_weight.  Golden from the unmodified reference
    (tests/golden/make_replay_golden.py, replay batch built by its GeneratedTerminationReplayDataset)."""
    z, meta, sd, grads = load_golden("replay_term")
    cfg = O.make_cfg(**meta["ctor"])
    idx, tgt = torch.from_numpy(z["idx"]), torch.from_numpy(z["targets"])
    replay = (torch.from_numpy(z["replay_x"]), torch.from_numpy(z["replay_labels"]))
    assert (replay[1] != -100).sum().item() > 0 and (replay[1] == -100).sum().item() > 0
    ow = {int(k): v for k, v in meta["offset_weights"].items()}
    total, parts, out, got = O.loss_and_grads(sd, cfg, idx, tgt, offset_weights=ow,
                                              termination_loss_weight=meta["termination_loss_weight"], replay=replay,
                                              replay_loss_weight=meta["replay_loss_weight"],
                                              replay_class_weights=torch.tensor(meta["replay_class_weights"]))
    assert parts["replay"].item() == pytest.approx(meta["parts"]["replay"], rel=3e-6)
    assert total.item() == pytest.approx(meta["parts"]["total"], rel=3e-6)
    gmax = max(v.norm().item() for v in grads.values())
    for k, ref in grads.items():  # key.bias gradients are analytically zero (pure rounding noise): absolute floor
        assert (got[k] - ref).norm().item() <= 2e-5 * max(ref.norm().item(), 1e-3 * gmax), k
