"""Time LayerNorm forward / backward at the C3 shape through the C ABI and report the HBM rate."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
from codonlm_b200 import ops  # noqa: E402

M, d = int(os.environ.get("LM", 65536)), int(os.environ.get("LD", 512))
dev = "cuda"
x = torch.randn(M, d, device=dev)
g = torch.randn(d, device=dev)
b = torch.randn(d, device=dev)
dy = torch.randn(M, d, device=dev).to(torch.bfloat16)
dres = torch.randn(M, d, device=dev)
dg, db, dxs = (torch.zeros(d, device=dev) for _ in range(3))
yb, _, mean, rstd = ops.layernorm_fwd(x, g, b)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def timeit(fn, nbytes, name):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    t = ts[len(ts) // 2]
    print(f"{name}: {t * 1e3:.1f} us  {nbytes / t / 1e6:.0f} GB/s (algorithmic bytes)")


timeit(lambda: ops.layernorm_fwd(x, g, b), M * d * 6, "ln_fwd (fp32 in, bf16 out)")
timeit(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dres, dg, db, want_bf16=True, dx_colsum=dxs), M * d * 16,
       "ln_bwd (bf16 dy, fp32 x, dres -> fp32 dx + bf16 copy + colsums)")
