"""Run attention forward on a few shapes and save (out, lse); with two files given, compare them.
  CGPT_ATTN_FWD=ws python tools/attn_fwd_compare.py run a.pt ; python tools/attn_fwd_compare.py run b.pt ;
  python tools/attn_fwd_compare.py cmp a.pt b.pt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))

SHAPES = [(2, 512, 4, 4, 64), (64, 1024, 8, 8, 64), (3, 333, 2, 2, 64), (2, 512, 8, 4, 48), (8, 4096, 8, 8, 64), (2, 200, 4, 2, 32)]

if sys.argv[1] == "run":
    from bench import synthetic_tokens
    from codonlm_b200 import ops
    res = {}
    for (B, T, H, Hk, hd) in SHAPES:
        g = torch.Generator().manual_seed(B * 1000 + T)
        qkv = torch.randn(B * T, (H + 2 * Hk) * hd, generator=g).to(torch.bfloat16).cuda()
        idx = synthetic_tokens(B, T, 77)[0].cuda()
        ss = ops.segment_starts(idx, 3)
        for name, s in (("seg", ss), ("causal", None)):
            out, lse = ops.attn_fwd(qkv, s, B, T, H, Hk, hd)
            torch.cuda.synchronize()
            res[(B, T, H, Hk, hd, name)] = (out.cpu(), lse.cpu())
    torch.save(res, sys.argv[2])
else:
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for k in a:
        oa, la = a[k]
        ob, lb = b[k]
        d = (oa.float() - ob.float()).abs()
        rel = d / oa.float().abs().clamp_min(1e-3)
        print(k, "out max abs diff %.3e" % d.max().item(), "elements differing %d of %d" % ((d > 0).sum().item(), d.numel()),
              "max rel %.3e" % rel.max().item(), "lse max diff %.3e" % (la - lb).abs().max().item())
