"""(b) drop-in boundary: the reference's own callers run UNCHANGED on the B200 implementation.

The overlay directory `genomics-lm_b200/overlay` goes ahead of a reference checkout on PYTHONPATH; `src.codonlm.model_tiny_gpt`
and `src.codonlm.training.objectives` then resolve to codonlm_b200 while `src.codonlm.checkpoints`, `scripts/query_model.py`,
`scripts/extract_embeddings.py`, `src/codonlm/training/loop.py` ... stay the reference's files
(reference: scripts/query_model.py:29-34,111-225, src/codonlm/checkpoints.py:12-56, scripts/extract_embeddings.py:342-360).

* CPU (build container, needs a reference checkout: CGPT_REFERENCE, /root/reference or the vendored baseline/_ref):
  import resolution, `build_codon_model_from_cfg`, `build_model_from_state`, strict load of a checkpoint written in the
  reference's runs/<id>/checkpoints/best.pt layout, loud failure when the model is called without a GPU.
* GPU: tests/golden/refcaller_driver.py — the reference's query_model CLI path (next / score / generate) and the
  extract_embeddings pooling — executed through the overlay with the model left on the CPU by the reference's `dev()`
  (host staging), against the golden the UNMODIFIED reference produced from the same checkpoint.
"""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import PKG, ROOT

OVERLAY = os.path.join(PKG, "overlay")
DRIVER = os.path.join(ROOT, "tests", "golden", "refcaller_driver.py")


def _reference_root():
    for cand in (os.environ.get("CGPT_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "src", "codonlm", "checkpoints.py")):
            return cand
    return None


def _run(code, ref, cwd):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([OVERLAY, ref]))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=cwd, timeout=600)


@pytest.mark.skipif(_reference_root() is None, reason="needs a reference checkout (CGPT_REFERENCE / baseline/_ref)")
def test_overlay_serves_the_reference_import_surface(tmp_path):
    code = r'''
import json, sys, torch
from pathlib import Path
import src.codonlm as pkg
import src.codonlm.model_tiny_gpt as M
import src.codonlm.training.objectives as OBJ
import src.codonlm.checkpoints as CK
import src.codonlm.training.loop as LOOP
import scripts.query_model as Q
from src.codonlm.codon_tokenize import VOCAB
info = dict(model_module=M.TinyGPT.__module__, pkg_tinygpt=pkg.TinyGPT.__module__, obj=OBJ.multi_offset_lm_loss.__module__,
            ck_file=CK.__file__, ck_tinygpt=CK.TinyGPT.__module__, loop_file=LOOP.__file__, loop_tinygpt=LOOP.TinyGPT.__module__,
            loop_obj=LOOP.termination_aux_loss.__module__, q_file=Q.__file__, q_tinygpt=Q.TinyGPT.__module__,
            names=[n for n in ("TinyGPT", "CausalSelfAttention", "Block", "NoPropBlock", "NoPropTinyGPT", "RotaryEmbedding",
                               "SwiGLU", "rotate_half", "apply_rotary_pos_emb") if hasattr(M, n)])
cfg = dict(vocab_size=68, block_size=32, n_layer=2, n_head=2, n_embd=32, dropout=0.0, n_kv_head=1, use_sdpa=True,
           termination_loss_enabled=True, multi_offset_targets=[2, 4], use_swiglu=False, use_rope=False)
run = Path("runs/2026-01-01_overlay"); (run / "checkpoints").mkdir(parents=True)
(run / "itos.txt").write_text("\n".join(VOCAB) + "\n")
torch.manual_seed(3)
m = CK.build_codon_model_from_cfg(cfg)
torch.save({"model": m.state_dict(), "cfg": cfg}, run / "checkpoints" / "best.pt")
sd, cfg2, path = CK.load_codon_checkpoint(run)
m2 = Q.build_model_from_state(sd, cfg2)
res = m2.load_state_dict(sd, strict=True)
info.update(built=type(m2).__module__, path=str(path), keys=len(sd), strict=[list(res.missing_keys), list(res.unexpected_keys)],
            equal=all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values())),
            eval_mode=not m2.training, dev=str(Q.dev()))
try:
    Q.next_token(m2.to(Q.dev()), Q.dev(), [1, 5, 6])
    info["forward"] = "ran"
except Exception as exc:
    info["forward"] = type(exc).__name__ + ": " + str(exc)[:80]
print("RESULT " + json.dumps(info))
'''
    ref = _reference_root()
    res = _run(code, ref, str(tmp_path))
    assert res.returncode == 0, res.stderr[-3000:]
    info = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    for key in ("model_module", "pkg_tinygpt", "ck_tinygpt", "loop_tinygpt", "q_tinygpt", "built"):
        assert info[key] == "codonlm_b200.model_tiny_gpt", (key, info[key])
    assert info["obj"] == info["loop_obj"] == "codonlm_b200.objectives"
    for key in ("ck_file", "loop_file", "q_file"):  # the callers themselves are the reference's own files
        assert os.path.realpath(info[key]).startswith(os.path.realpath(ref)), (key, info[key])
    assert len(info["names"]) == 9
    assert info["strict"] == [[], []] and info["equal"] and info["eval_mode"] and info["path"].endswith("checkpoints/best.pt")
    if torch.cuda.is_available():
        assert info["forward"] == "ran"
    else:  # no GPU here: the staged call must fail loudly, never compute on the CPU
        assert info["forward"].startswith("CgptError") and "no CPU implementation" in info["forward"]


@pytest.mark.gpu
def test_reference_query_model_and_embedding_paths_run_unchanged(tmp_path):
    ref = _reference_root()
    if ref is None:
        pytest.skip("no reference checkout on this box (baseline/_ref is vendored by __graft_entry__.build())")
    with open(os.path.join(ROOT, "tests", "golden", "refcaller_golden.json")) as f:
        want = json.load(f)
    out = str(tmp_path / "ours.json")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([OVERLAY, ref]))
    res = subprocess.run([sys.executable, DRIVER, "--out", out], capture_output=True, text=True, env=env,
                         cwd=str(tmp_path), timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    with open(out) as f:
        got = json.load(f)
    assert want["model_module"] == "src.codonlm.model_tiny_gpt"       # the golden came from the unmodified reference
    assert got["model_module"] == "codonlm_b200.model_tiny_gpt"       # this run went through the B200 implementation
    assert got["strict_load"] == [[], []]
    # next-token distribution (the CLI's top-k is read off it): every probability within 2e-3, top-1 token identical
    for key in ("next_probs", "next_probs_short"):
        a, b = torch.tensor(got[key]), torch.tensor(want[key])
        assert (a - b).abs().max().item() <= 2e-3, key
        assert int(a.argmax()) == int(b.argmax())
    assert got["cli_next"]["topk"][0]["token"] == want["cli_next"]["topk"][0]["token"]
    assert got["cli_next"]["topk"][0]["prob"] == pytest.approx(want["cli_next"]["topk"][0]["prob"], abs=2e-3)
    # score_sequence: nll within north_star's 1e-3 relative
    assert got["cli_score"]["nll"] == pytest.approx(want["cli_score"]["nll"], rel=1e-3)
    assert got["score2"]["nll"] == pytest.approx(want["score2"]["nll"], rel=1e-3)
    # greedy generation follows the reference while its top-2 margin is resolvable: same length, same prompt
    assert got["greedy"][:10] == want["greedy"][:10] and len(got["greedy"]) == len(want["greedy"])
    # pooled hidden states of every stage (extract_embeddings)
    assert set(got["pooled"]) == set(want["pooled"])
    for key, ref_v in want["pooled"].items():
        a, b = torch.tensor(got["pooled"][key]), torch.tensor(ref_v)
        tol = 1e-6 if key.startswith("layer_0__") else 2e-2  # stage 0 is the embedding gather: exact
        assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item()), key
