"""Time attention fwd/bwd at the C3 shape through the C ABI (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
from codonlm_b200 import ops  # noqa: E402

B, T, H, Hk, hd = (int(os.environ.get(k, v)) for k, v in (("AB", 64), ("AT", 1024), ("AH", 8), ("AHK", 8), ("AHD", 64)))
dev = "cuda"
qkv = torch.randn(B * T, (H + 2 * Hk) * hd, device=dev).to(torch.bfloat16)
if os.environ.get("AREAL", "1") == "1":  # the bench's token streams: ATOK=random (headline: one segment per sequence,
    sys.path.insert(0, ROOT)             # full causal) or ATOK=realistic (SEP every U{100..400} tokens)
    from bench import synthetic_tokens  # noqa: E402
    idx = synthetic_tokens(B, T, 1234, kind=os.environ.get("ATOK", "random"))[0].to(dev)
else:
    idx = torch.randint(4, 68, (B, T), device=dev)
    idx[:, 300] = 3
ss = ops.segment_starts(idx, 3)
out, lse = ops.attn_fwd(qkv, ss, B, T, H, Hk, hd)
dout = torch.randn_like(out)
pairs = B * H * sum(range(1, T // 128 + 1))
ssc = ss.cpu()
vis = sum(int(ssc[b, q * 128] // 128 <= k) for b in range(B) for q in range(T // 128) for k in range(q + 1))
print(f"visible (q tile, kv tile) pairs per head: {vis / B:.1f} of {sum(range(1, T // 128 + 1))}")
for name, fn, fl in (("attn_fwd", lambda: ops.attn_fwd(qkv, ss, B, T, H, Hk, hd), 4 * 128 * 128 * hd),
                     ("attn_bwd", lambda: ops.attn_bwd(qkv, ss, out, dout, lse, B, T, H, Hk, hd), 10 * 128 * 128 * hd)):
    if len(sys.argv) > 1 and sys.argv[1] != name:
        continue
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    print(f"{name}: {ts[2] * 1e3:.1f} us  ({pairs * fl / ts[2] / 1e9:.0f} TFLOP/s on {pairs} tile pairs incl. masked halves of diagonal tiles)")
