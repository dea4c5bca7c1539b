"""Time individual GEMM shapes of the step through the C ABI (CUDA events), for ncu captures.
usage: python tools/gemm_probe.py [shape_name ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
from codonlm_b200 import ops  # noqa: E402

M = 65536
SHAPES = {
    # name: (M, N, K, a_mn, b_mn, out dtype, epilogue, residual, split)
    "qkv_fwd": (M, 1536, 512, False, False, torch.bfloat16, ops.EPI_NONE, False, 1),
    "fc1_fwd_gelu": (M, 2048, 512, False, False, torch.bfloat16, ops.EPI_GELU, False, 1),
    "fc2_dgrad_mulaux": (M, 2048, 512, False, True, torch.bfloat16, ops.EPI_MUL_AUX, False, 1),
    "fc2_dgrad_mulaux_colsum": (M, 2048, 512, False, True, torch.bfloat16, ops.EPI_MUL_AUX, False, 1),
    "proj_fwd_res": (M, 512, 512, False, False, torch.float32, ops.EPI_NONE, True, 1),
    "fc2_fwd_res": (M, 512, 2048, False, False, torch.float32, ops.EPI_NONE, True, 1),
    "fc1_dgrad": (M, 512, 2048, False, True, torch.bfloat16, ops.EPI_NONE, False, 1),
    "fc1_wgrad": (2048, 512, M, True, True, torch.float32, ops.EPI_NONE, False, 23),
    "proj_wgrad": (512, 512, M, True, True, torch.float32, ops.EPI_NONE, False, 18),
}


def run(name, iters=10):
    m, n, k, amn, bmn, dt, epi, res, split = SHAPES[name]
    dev = "cuda"
    a = torch.randn((k, m) if amn else (m, k), device=dev).to(torch.bfloat16)
    b = torch.randn((k, n) if bmn else (n, k), device=dev).to(torch.bfloat16)
    out = torch.zeros((m, n), dtype=dt, device=dev)
    bias = torch.randn(n, device=dev) if split == 1 else None
    aux = torch.randn((m, n), device=dev).to(torch.bfloat16) if epi != ops.EPI_NONE else None
    resid = torch.randn((m, n), device=dev) if res else None
    kw = dict(M=m, N=n, K=k, a_mn=amn, b_mn=bmn, bias=bias, epilogue=epi, residual=resid, split_k=split,
              accumulate=split > 1)
    if epi == ops.EPI_GELU:
        kw.update(aux_out=aux, ldaux=n)
    elif epi == ops.EPI_MUL_AUX:
        kw.update(aux=aux, ldaux=n)
        if name.endswith("_colsum"):
            kw.update(colsum=torch.zeros(n, device=dev))
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)  # 256 MB > L2
    for _ in range(2):
        ops.gemm(a, b, out, **kw)
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.gemm(a, b, out, **kw)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:18s} M{m} N{n} K{k}: {med * 1e3:8.1f} us  {2.0 * m * n * k / med / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(SHAPES)):
        run(nm)
