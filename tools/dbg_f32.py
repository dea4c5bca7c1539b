import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "genomics-lm_b200"))
from codonlm_b200 import ops, TinyGPT
torch.manual_seed(0)
for (M, N, K) in [(256, 64, 256), (256, 64, 176), (512, 512, 512)]:
    A = (torch.randn(M, K) * 0.5).to(torch.bfloat16).cuda(); W = (torch.randn(N, K) * 0.1).to(torch.bfloat16).cuda()
    b = torch.randn(N).cuda()
    out = torch.full((M, N), 777.0, device="cuda")
    ops.gemm(A, W, out, M=M, N=N, K=K, bias=b)
    ref = A.float() @ W.float().t() + b
    print("plain f32+bias", M, N, K, (out - ref).abs().max().item())
    res = torch.randn(M, N).cuda()
    ops.gemm(A, W, out, M=M, N=N, K=K, bias=b, residual=res)
    print("res   f32+bias", M, N, K, (out - ref - res).abs().max().item())
m = TinyGPT(vocab_size=68, block_size=64, n_layer=2, n_head=2, n_embd=64, dropout=0.1, use_sdpa=True).cuda()
x = torch.randint(4, 68, (4, 64), device="cuda")
m.train()
for i in range(3):
    lg, loss = m(x, x)
    print("train loss", loss.item(), lg.abs().max().item(), torch.isfinite(lg).all().item())
    loss.backward()
opt = torch.optim.AdamW(m.parameters(), lr=3e-3)
for i in range(3):
    lg, loss = m(x, x); opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
    print("opt loop loss", loss.item(), lg.abs().max().item())
