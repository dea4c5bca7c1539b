"""Drives the REFERENCE'S OWN inference callers end to end and prints what they return as JSON:

  * writes a run directory in the reference's layout (runs/<RUN_ID>/checkpoints/best.pt = {"model", "cfg"}, itos.txt),
  * `scripts.query_model.run_once` (the CLI path: _load_vocab, load_codon_checkpoint, build_model_from_state,
    model.to(dev()), next_token / score_sequence) in modes "next" and "score", plus `next_token` distributions and a
    greedy `generate`,
  * `model.iter_hidden_states` pooled by `scripts.extract_embeddings._pool_state` (the extract_embeddings path).

Every module it imports is resolved through sys.path, so the SAME file runs (a) against the unmodified reference
(PYTHONPATH=<reference>): that output is committed as tests/golden/refcaller_golden.json, and (b) with this repo's
overlay ahead of the (vendored) reference (PYTHONPATH=genomics-lm_b200/overlay:baseline/_ref), where
`src.codonlm.model_tiny_gpt` is the B200 implementation and the callers are the reference's unchanged files.

    PYTHONPATH=/root/reference python tests/golden/refcaller_driver.py --out tests/golden/refcaller_golden.json
"""
import argparse
import json
import os
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import torch

CFG = dict(vocab_size=68, block_size=128, n_layer=2, n_head=4, n_embd=128, dropout=0.0, label_smoothing=0.0,
           sep_mask_enabled=True, tie_embeddings=True, use_sdpa=True, n_kv_head=2)
DNA = "ATGGCTAAAGGTCTGACCGAATTTGCAGGCCGTATCGTTAACCTGGAAGATCTGAAAGCGTTTCGTGAACATCCGGGCTGA"
DNA2 = "ATGACCATGATTACGCCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTCTAA"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    import scripts.query_model as Q
    from scripts.extract_embeddings import _pool_state
    from src.codonlm.checkpoints import build_codon_model_from_cfg, load_codon_checkpoint
    from src.codonlm.codon_tokenize import VOCAB

    out = {"model_module": None}
    with tempfile.TemporaryDirectory() as td:
        run_dir = Path(td) / "runs" / "2026-01-01_refcaller"
        (run_dir / "checkpoints").mkdir(parents=True)
        (run_dir / "itos.txt").write_text("\n".join(VOCAB) + "\n")
        torch.manual_seed(1337)  # the constructor's RNG contract makes these bits identical in both implementations
        m0 = build_codon_model_from_cfg(CFG)
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            m0.tok_emb.weight.mul_(0.05)
            m0.pos_emb.weight.mul_(0.05)
            for name, p in m0.named_parameters():
                if ".ln" in name or name.startswith("ln_f"):
                    p.add_(0.1 * torch.randn(p.shape, generator=g))
        torch.save({"model": m0.state_dict(), "cfg": dict(CFG)}, run_dir / "checkpoints" / "best.pt")

        sd, cfg, path = load_codon_checkpoint(run_dir)
        assert path.name == "best.pt" and cfg["n_embd"] == CFG["n_embd"]
        model = Q.build_model_from_state(sd, cfg)
        missing = model.load_state_dict(sd, strict=True)  # strict: the state-dict layout is the checkpoint contract
        out["model_module"] = type(model).__module__
        out["strict_load"] = [list(missing.missing_keys), list(missing.unexpected_keys)]
        cli = dict(run_id=None, run_dir=str(run_dir), dna=DNA, interactive=False, topk=5, temperature=1.0, max_new=8)
        out["cli_next"] = Q.run_once(SimpleNamespace(mode="next", **cli))
        out["cli_score"] = Q.run_once(SimpleNamespace(mode="score", **cli))
        device = Q.dev()
        model.to(device)
        itos, stoi = Q._load_vocab(run_dir)
        ids = Q.dna_to_ids(DNA, stoi)
        ids2 = Q.dna_to_ids(DNA2, stoi)
        out["next_probs"] = torch.softmax(Q.next_token(model, device, ids), dim=-1).tolist()
        out["next_probs_short"] = torch.softmax(Q.next_token(model, device, ids[:5]), dim=-1).tolist()
        out["score2"] = Q.score_sequence(model, device, ids2)
        out["greedy"] = Q.generate(model, device, ids[:10], max_new=6, topk=1)
        # extract_embeddings path: batch of two padded sequences, every stage, mean over non-pad positions + eos state
        T = max(len(ids), len(ids2))
        x = torch.zeros((2, T), dtype=torch.long, device=device)
        x[0, : len(ids)] = torch.tensor(ids)
        x[1, : len(ids2)] = torch.tensor(ids2)
        nonpad = x.ne(0)
        pooled = {}
        with torch.no_grad():
            for layer, hidden in model.iter_hidden_states(x):
                for mode in ("mean_nonpad", "eos"):
                    pooled[f"layer_{layer}__{mode}"] = _pool_state(hidden, x, nonpad, mode=mode, content_ids=set()).tolist()
        out["pooled"] = pooled
    with open(args.out, "w") as f:
        json.dump(out, f)
    print("model module:", out["model_module"], "| nll:", out["cli_score"]["nll"])


if __name__ == "__main__":
    sys.path.insert(0, os.getcwd())
    main()
