"""CPU checks of the drop-in boundary: constructor, state_dict keys/shapes and
bit-identical initial weights against what the reference produced (tests/golden meta), errors."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden


def _sha(model):
    h = hashlib.sha256()
    for k, v in model.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_state_dict_layout_and_init_bits_match_reference(case):
    from codonlm_b200 import TinyGPT
    z, meta, sd, grads = load_golden(case)
    torch.manual_seed(1337)
    m = TinyGPT(**meta["ctor"])
    mine = {k: v for k, v in m.state_dict().items() if not k.endswith("attn.mask")}
    assert list(mine) == list(sd), "state_dict key order differs from the reference"
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
    # same construction order => same RNG stream => same bits as the reference constructor
    assert _sha(m) == meta["init_sha256"]
    # strict load of a reference checkpoint (loop.py:882 resumes strictly)
    full = dict(sd)
    for l in range(meta["ctor"]["n_layer"]):
        bs = meta["ctor"]["block_size"]
        full[f"blocks.{l}.attn.mask"] = torch.tril(torch.ones(bs, bs)).view(1, 1, bs, bs)
    m.load_state_dict(full, strict=True)
    assert set(n for n, _ in m.named_parameters()) == set(grads) | ({"head.weight"} - set(grads)) - (
        {"head.weight"} if meta["ctor"].get("tie_embeddings", True) else set())


def test_constructor_contract():
    from codonlm_b200 import TinyGPT
    m = TinyGPT(69, 16, n_layer=1, n_head=4, n_embd=32, n_kv_head=0)
    assert m.n_kv_head is None and m.blocks[0].attn.n_kv_head is None
    with pytest.raises(AssertionError):
        TinyGPT(69, 16, n_layer=1, n_head=3, n_embd=32)
    d = TinyGPT(68, 8, n_layer=1, n_head=1, n_embd=16, multi_offset_targets=[4, 2, 2], termination_aux=True).to_dict()
    assert d["multi_offset_targets"] == [2, 4] and d["termination_aux"] and d["sep_mask_enabled"]
    assert m.head.weight is m.tok_emb.weight
    assert TinyGPT(68, 8, n_layer=1, n_head=1, n_embd=16, use_rope=True).pos_emb is None


def test_mask_truth_table_on_module():  # reference tests/test_models.py:29-51
    from codonlm_b200 import TinyGPT
    model = TinyGPT(vocab_size=8, block_size=5, n_layer=1, n_head=1, n_embd=8, dropout=0.0, sep_id=3)
    tokens = torch.tensor([[1, 4, 3, 5, 6]])
    full = model.build_attention_mask(tokens)[0, 0]
    assert full[1, 0] and not full[3, 1] and full[3, 2] and full[4, 2]
    local = model.build_attention_mask(tokens, attention_window=1)[0, 0]
    assert torch.equal(local, torch.eye(5, dtype=torch.bool))
    with pytest.raises(ValueError, match="at least 1"):
        model.build_attention_mask(tokens, attention_window=0)


def test_cpu_forward_fails_loudly():
    from codonlm_b200 import TinyGPT, _lib
    m = TinyGPT(68, 8, n_layer=1, n_head=1, n_embd=16, dropout=0.0)
    with pytest.raises(_lib.CgptError, match="no CPU implementation"):
        m(torch.zeros((1, 4), dtype=torch.long))
    m.eval()  # host-resident inference callers are staged to a B200 — which this box does not have: still loud
    with torch.no_grad(), pytest.raises(_lib.CgptError, match="no CPU implementation"):
        m(torch.zeros((1, 4), dtype=torch.long))
    with torch.no_grad(), pytest.raises(_lib.CgptError, match="no CPU implementation"):
        m.forward_hidden(torch.zeros((1, 4), dtype=torch.long))
    with torch.no_grad(), pytest.raises(_lib.CgptError, match="no CPU implementation"):
        m.blocks[0].attn(torch.zeros((1, 4, 16)))


def test_boolean_masks_convert_to_interval_starts():
    """mask_spec_from_bool (the reference's boolean-mask calling convention, model_tiny_gpt.py:106-113): every mask the
    oracle's build_attention_mask restatement produces maps to first-visible-column = max(segment start, i-window+1);
    masks that are not causal intervals are refused.  Pure index logic: runs on the CPU."""
    import numpy as np
    import torch
    from codonlm_b200.model_tiny_gpt import mask_spec_from_bool
    from oracle import codon_gpt_oracle as O
    idx, _ = O.synthetic_batch(3, 50, seed=8, realistic=True)
    idx[:, 13], idx[2, 31] = 3, 3
    for window in (None, 1, 6):
        m = torch.from_numpy(np.ascontiguousarray(O.attention_mask(idx.numpy(), 3, window)))
        spec = mask_spec_from_bool(m, 3, 50)
        seg_start = torch.zeros_like(idx)
        for b in range(3):
            last = 0
            for t in range(50):
                if int(idx[b, t]) == 3:
                    last = t
                seg_start[b, t] = last
        want = seg_start if window is None else torch.maximum(seg_start, torch.arange(50)[None] - window + 1)
        assert spec.window == 0 and torch.equal(spec.seg_start.long(), want)
    tri = torch.tril(torch.ones(50, 50, dtype=torch.bool))
    assert torch.equal(mask_spec_from_bool(tri, 3, 50).seg_start, torch.zeros((3, 50), dtype=torch.int32))
    assert torch.equal(mask_spec_from_bool(tri[None], 3, 50).seg_start, torch.zeros((3, 50), dtype=torch.int32))
    holed = tri.clone()
    holed[9, 4] = False
    full = torch.ones(50, 50, dtype=torch.bool)  # not causal
    for bad in (holed, full, tri[None, None].expand(3, 2, 50, 50)):
        with pytest.raises(NotImplementedError):
            mask_spec_from_bool(bad, 3, 50)


def test_gradient_by_products_travel_on_the_tensor_and_through_views():
    """The bf16 form of a logit gradient that the cross-entropy backward writes (functional.CrossEntropyFn) reaches the
    head's backward ON the gradient tensor — also through the (B*T,V) <-> (B,T,V) views autograd puts in between —
    is consumed once, and is ignored when its shape is not the one the consumer wants (it then recomputes)."""
    from codonlm_b200 import functional as Fn
    g = torch.zeros(12, 68)
    side = torch.zeros(12, 72, dtype=torch.bfloat16)
    Fn._attach(g, side, None)
    via_view = g.view(3, 4, 68).view(12, 68)  # what ViewBackward hands on
    assert Fn._grad_form_of(via_view, (12, 3 * 72)) is None  # not the form this consumer reads: it will recompute
    assert Fn._grad_form_of(via_view, (12, 72)) is None      # ... and the by-product is gone: it is looked up once
    Fn._attach(g, side, None)
    assert Fn._grad_form_of(g.view(3, 4, 68).view(12, 68), (12, 72)) is side
    assert Fn._grad_form_of(g, (12, 72)) is None
    # a tensor that is not a full view of the producer's (a slice, a sum) never sees it
    Fn._attach(g, side, None)
    assert Fn._grad_form_of(g[:6], (6, 72)) is None
    assert Fn._grad_form_of(g + 0, (12, 72)) is None
    assert Fn._grad_form_of(g, (12, 72)) is side


def test_multi_output_nodes_take_undefined_gradients_as_none():
    """ResidualLayerNormFn / HeadsFn / CrossEntropyFn switch gradient materialisation off (an unused output must not cost
    a zero fill, a conversion and an add in backward): checked on the autograd metadata, no kernel is launched."""
    import inspect
    from codonlm_b200 import functional as Fn
    for cls in (Fn.ResidualLayerNormFn, Fn.HeadsFn, Fn.CrossEntropyFn):
        assert "set_materialize_grads(False)" in inspect.getsource(cls.forward), cls.__name__
