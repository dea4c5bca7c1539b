"""Per-kernel parity tests (B200 only): each C-ABI entry point against a plain PyTorch fp32 restatement
of the same arithmetic on the same seeded inputs.  Integer kernels are bit-exact; bf16 GEMMs are
bit-exact on small-integer operands and toleranced on random ones."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import codon_gpt_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from codonlm_b200 import ops as _ops
    return _ops


DEV = "cuda"


def _randint_bf16(shape, g, lo=-2, hi=3):
    return torch.randint(lo, hi, shape, generator=g).to(torch.bfloat16).to(DEV)


# ------------------------------------------------------------------------------------------ scans
def test_integer_scans_bit_exact(ops):
    rng = np.random.default_rng(0)
    for (B, T) in [(1, 1), (3, 31), (4, 257), (8, 1024), (2, 4096)]:
        idx = rng.integers(0, 12, size=(B, T), dtype=np.int64)
        t = torch.from_numpy(idx).to(DEV)
        seg = ops.segment_ids(t, 3).cpu().numpy()
        assert np.array_equal(seg, O.segment_ids(idx, 3))
        st = ops.segment_starts(t, 3).cpu().numpy()
        exp = np.zeros_like(idx)
        for b in range(B):
            cur = 0
            for j in range(T):
                if idx[b, j] == 3:
                    cur = j
                exp[b, j] = cur
        assert np.array_equal(st, exp)
        for stops, edges in (((2,), (0, 3, 10, 30)), ((2, 3), (0, 1, 3)), ((2,), ())):
            nxt = ops.next_in_set(t, stops)
            lab = ops.termination_labels(t, nxt, edges).cpu().numpy()
            assert np.array_equal(lab, O.termination_distance_bucket_labels(idx, stops, edges))


def test_scans_match_reference_vectors(ops):
    from conftest import GOLDEN_DIR
    z = np.load(f"{GOLDEN_DIR}/integer_kats.npz")
    yb = torch.from_numpy(z["yb"]).to(DEV)
    for name, (stops, edges) in {"a": ((2,), (0, 3, 10, 30)), "b": ((2, 3), (0, 1, 3)), "c": ((2,), ())}.items():
        lab = ops.termination_labels(yb, ops.next_in_set(yb, stops), edges).cpu().numpy()
        assert np.array_equal(lab, z[f"term.{name}"])
    nb = ops.next_in_set(yb, (2, 3)).cpu().numpy()
    B, T = z["yb"].shape
    for o in (2, 3, 4, 8, 16, 32):
        ref = z[f"offset_mask.{o}"]
        t = np.arange(T - o + 1)[None, :]
        mine = (z["yb"][:, o - 1:] != 0) & (nb[:, : T - o + 1] >= t + o - 1)
        assert np.array_equal(mine, ref), o


# ------------------------------------------------------------------------------------------ embedding / LN
def test_embed_fwd_bwd(ops):
    g = torch.Generator().manual_seed(1)
    for (B, T, d, V, pos) in [(2, 16, 32, 68, True), (4, 128, 512, 68, True), (3, 50, 256, 69, False)]:
        idx = torch.randint(0, V, (B, T), generator=g).to(DEV)
        tok = torch.randn(V, d, generator=g).to(DEV)
        pw = torch.randn(T + 3, d, generator=g).to(DEV) if pos else None
        x = ops.embed_fwd(idx, tok, pw)
        ref = tok[idx] + (pw[:T][None] if pos else 0)
        assert torch.equal(x, ref)
        dx = torch.randn(B, T, d, generator=g).to(DEV)
        dtok = torch.zeros_like(tok)
        dpos = torch.zeros_like(pw) if pos else None
        ops.embed_bwd(idx, dx, dtok, dpos)
        rt = torch.zeros_like(tok).index_add_(0, idx.reshape(-1), dx.reshape(-1, d))
        assert torch.allclose(dtok, rt, rtol=1e-5, atol=1e-4)
        if pos:
            rp = torch.zeros_like(pw)
            rp[:T] = dx.sum(0)
            assert torch.allclose(dpos, rp, rtol=1e-5, atol=1e-4)


def test_layernorm_fwd_bwd(ops):
    g = torch.Generator().manual_seed(2)
    # (1024, 512) and (20011, 512) run the streaming kernels (bulk-copy row rings; the second one re-uses every ring slot
    # several times and ends on ragged per-warp row counts)
    for (M, d) in [(7, 32), (300, 64), (1024, 512), (513, 384), (64, 1024), (20011, 512)]:
        x = (torch.randn(M, d, generator=g) * 2 + 0.5).to(DEV)
        gam = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV)
        bet = (0.1 * torch.randn(d, generator=g)).to(DEV)
        yb, yf, mean, rstd = ops.layernorm_fwd(x, gam, bet, want_bf16=True, want_f32=True)
        ref = torch.nn.functional.layer_norm(x, (d,), gam, bet, 1e-5)
        assert torch.allclose(yf, ref, rtol=1e-5, atol=2e-6)
        assert torch.equal(yb, yf.to(torch.bfloat16))
        dy = torch.randn(M, d, generator=g).to(DEV)
        dres = torch.randn(M, d, generator=g).to(DEV)
        xr = x.clone().requires_grad_(True)
        gr = gam.clone().requires_grad_(True)
        br = bet.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy)
        for dyt in (dy, dy.to(torch.bfloat16)):
            dgam, dbet = torch.zeros_like(gam), torch.zeros_like(bet)
            dxs = torch.ones_like(gam)
            dx, dxb = ops.layernorm_bwd(dyt, x, gam, mean, rstd, dres, dgam, dbet, want_bf16=True, dx_colsum=dxs)
            assert torch.allclose(dxs, 1 + dx.sum(0), rtol=1e-4, atol=1e-3 * math.sqrt(M))
            tol = 1e-4 if dyt.dtype == torch.float32 else 2e-2
            assert torch.allclose(dx, xr.grad + dres, rtol=tol, atol=tol)
            assert torch.allclose(dgam, gr.grad, rtol=tol, atol=tol * math.sqrt(M))
            assert torch.allclose(dbet, br.grad, rtol=tol, atol=tol * math.sqrt(M))
            assert torch.equal(dxb, dx.to(torch.bfloat16))
        # no incoming residual gradient, no bf16 copy, no column sums
        dgam, dbet = torch.zeros_like(gam), torch.zeros_like(bet)
        dx, dxb = ops.layernorm_bwd(dy.to(torch.bfloat16), x, gam, mean, rstd, None, dgam, dbet)
        assert dxb is None
        assert torch.allclose(dx, xr.grad, rtol=2e-2, atol=2e-2)
        assert torch.allclose(dgam, gr.grad, rtol=2e-2, atol=2e-2 * math.sqrt(M))


# ------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [(128, 256, 64), (128, 64, 64), (256, 512, 512), (300, 200, 136), (1000, 1536, 512), (64, 72, 1368),
               (1, 256, 64), (8, 1152, 384)]  # decode-sized M


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_layouts_exact(ops, a_mn, b_mn):
    g = torch.Generator().manual_seed(3)
    for (M, N, K) in GEMM_SHAPES:
        A = _randint_bf16((M, K), g)
        Bm = _randint_bf16((N, K), g)
        ref = A.float() @ Bm.float().t()
        # MN-major operands need a 16-byte pitch along M / N: pad the stored pitch to a multiple of 8
        def store(mat, mn):
            if not mn:
                return mat.contiguous(), mat.shape[1]
            rows = mat.shape[0]
            ld = (rows + 7) // 8 * 8
            buf = torch.zeros((mat.shape[1], ld), dtype=torch.bfloat16, device=DEV)
            buf[:, :rows] = mat.t()
            return buf, ld
        a, lda = store(A, a_mn)
        b, ldb = store(Bm, b_mn)
        for dt in (torch.float32, torch.bfloat16):
            out = torch.full((M, N), 7.0, dtype=dt, device=DEV)
            ops.gemm(a, b, out, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, lda=lda, ldb=ldb)
            torch.cuda.synchronize()
            exp = ref if dt == torch.float32 else ref.to(torch.bfloat16)
            bad = (out != exp)
            assert not bad.any(), (f"M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn} {dt}: {int(bad.sum())} wrong, first at "
                                   f"{bad.nonzero()[0].tolist()}, got {out[bad][:4].tolist()} want {exp[bad][:4].tolist()}")


def test_gemm_epilogues(ops):
    g = torch.Generator().manual_seed(4)
    M, N, K = 520, 384, 256
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    ref = A.float() @ W.float().t() + bias
    # bias + GELU; the side output is gelu'(pre) for the backward epilogue
    out = torch.empty((M, N), dtype=torch.bfloat16, device=DEV)
    dact = torch.empty((M, N), dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, out, M=M, N=N, K=K, bias=bias, epilogue=ops.EPI_GELU, aux_out=dact, ldaux=N)
    p32 = ref.clone().requires_grad_(True)
    gel = torch.nn.functional.gelu(p32)
    gel.sum().backward()
    assert torch.allclose(out.float(), gel.detach(), rtol=1e-2, atol=1e-2)
    assert torch.allclose(dact.float(), p32.grad, rtol=1e-2, atol=1e-2)
    # fast erf: check the fused activation at fp32 precision through an fp32-representable path
    xs = torch.linspace(-8, 8, 4096, device=DEV)
    eye = torch.zeros((4096, 8), dtype=torch.bfloat16, device=DEV)
    eye[:, 0] = 1
    wx = torch.zeros((8, 8), dtype=torch.bfloat16, device=DEV)
    wx[0, 0] = 1
    bvec = torch.zeros(8, device=DEV)
    for shift in (0.0, 0.00390625):  # exact bf16-representable grid points as bias, value through the GEMM
        col = torch.zeros((4096, 8), dtype=torch.bfloat16, device=DEV)
        col[:, 0] = xs.to(torch.bfloat16)
        o2 = torch.empty((4096, 8), dtype=torch.bfloat16, device=DEV)
        d2 = torch.empty((4096, 8), dtype=torch.bfloat16, device=DEV)
        bvec[0] = shift
        ops.gemm(col, wx, o2, M=4096, N=8, K=8, bias=bvec, epilogue=ops.EPI_GELU, aux_out=d2, ldaux=8)
        xv = (col[:, 0].float() + shift).requires_grad_(True)
        gv = torch.nn.functional.gelu(xv)
        gv.sum().backward()
        assert torch.equal(o2[:, 0], gv.detach().to(torch.bfloat16)) or \
            (o2[:, 0].float() - gv.detach()).abs().max().item() <= 2 ** -8 * gv.detach().abs().max().item()
        assert (d2[:, 0].float() - xv.grad).abs().max().item() <= 2 ** -7
    # ragged GELU tile: N not a multiple of 8, odd aux pitch -> scalar tail path
    Nr = 100
    outr = torch.empty((M, Nr), dtype=torch.bfloat16, device=DEV)
    dr = torch.empty((M, Nr), dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W[:Nr].contiguous(), outr, M=M, N=Nr, K=K, bias=bias[:Nr].contiguous(), epilogue=ops.EPI_GELU,
             aux_out=dr, ldaux=Nr)
    assert torch.allclose(outr.float(), gel.detach()[:, :Nr], rtol=1e-2, atol=1e-2)
    assert torch.allclose(dr.float(), p32.grad[:, :Nr], rtol=1e-2, atol=1e-2)
    # multiply-by-aux epilogue (backward through GELU)
    dz = torch.empty((M, N), dtype=torch.bfloat16, device=DEV)
    cs = torch.ones(N, device=DEV)
    ops.gemm(A, W, dz, M=M, N=N, K=K, epilogue=ops.EPI_MUL_AUX, aux=dact, ldaux=N, colsum=cs)
    assert torch.allclose(dz.float(), (A.float() @ W.float().t()) * dact.float(), rtol=2e-2, atol=2e-2)
    assert torch.allclose(cs, 1 + dz.float().sum(0), rtol=1e-4, atol=1e-3)  # fused column sums of the bf16 output
    # bias + residual -> fp32
    res = torch.randn(M, N, generator=g).to(DEV)
    o32 = torch.empty((M, N), dtype=torch.float32, device=DEV)
    ops.gemm(A, W, o32, M=M, N=N, K=K, bias=bias, residual=res)
    assert torch.allclose(o32, ref + res, rtol=1e-4, atol=1e-3)
    # split-K accumulate (wgrad form): dW[N,K] += dYᵀ X over M tokens
    Mt = 4096
    dY = (torch.randn(Mt, N, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    X = (torch.randn(Mt, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    dW = torch.ones((N, K), dtype=torch.float32, device=DEV)
    ops.gemm(dY, X, dW, M=N, N=K, K=Mt, a_mn=True, b_mn=True, accumulate=True, split_k=8)
    assert torch.allclose(dW, 1 + dY.float().t() @ X.float(), rtol=1e-3, atol=1e-2)
    # odd output pitch (SwiGLU hidden 1365-like): fp32 accumulate into [N, 85]
    Kodd = 85
    X2 = (torch.randn(Mt, 88, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    X2[:, Kodd:] = 0
    dW2 = torch.zeros((N, Kodd), dtype=torch.float32, device=DEV)
    ops.gemm(dY, X2, dW2, M=N, N=Kodd, K=Mt, a_mn=True, b_mn=True, ldb=88, accumulate=True, split_k=4)
    assert torch.allclose(dW2, dY.float().t() @ X2[:, :Kodd].float(), rtol=1e-3, atol=1e-2)


# ------------------------------------------------------------------------------------------ small elementwise
def test_cast_colsum_rope_swiglu_adamw(ops):
    g = torch.Generator().manual_seed(5)
    w = torch.randn(85, 64, generator=g).to(DEV)
    o = ops.cast_bf16(w, ld_out=72)
    assert torch.equal(o[:, :64], w.to(torch.bfloat16)) and not o[:, 64:].any()
    assert torch.equal(ops.cast_bf16(w), w.to(torch.bfloat16))
    x = torch.randn(1000, 200, generator=g).to(torch.bfloat16).to(DEV)
    cs = torch.ones(200, device=DEV)
    ops.colsum_bf16(x, cs)
    assert torch.allclose(cs, 1 + x.float().sum(0), rtol=1e-4, atol=1e-3)
    # rope
    B, T, H, Hk, hd = 2, 40, 4, 2, 32
    qkv = torch.randn(B * T, (H + 2 * Hk) * hd, generator=g).to(torch.bfloat16).to(DEV)
    cos, sin = O.rope_tables(T, hd)
    c2, s2 = cos[:, : hd // 2].contiguous().to(DEV), sin[:, : hd // 2].contiguous().to(DEV)
    ref = qkv.float().view(B, T, H + 2 * Hk, hd).clone()
    qk = ref[:, :, : H + Hk].permute(0, 2, 1, 3)
    ref[:, :, : H + Hk] = O.apply_rope(qk, cos.to(DEV), sin.to(DEV)).permute(0, 2, 1, 3)
    mine = qkv.clone()
    ops.rope_qk(mine, c2, s2, B, T, H, Hk, hd)
    assert torch.allclose(mine.float().view(B, T, -1, hd), ref, rtol=1e-2, atol=1e-2)
    ops.rope_qk(mine, c2, s2, B, T, H, Hk, hd, inverse=True)
    assert torch.allclose(mine.float(), qkv.float(), rtol=2e-2, atol=2e-2)
    # swiglu
    M, hp = 300, 88
    gu = torch.randn(M, 2 * hp, generator=g).to(torch.bfloat16).to(DEV)
    act = ops.swiglu_fwd(gu, hp)
    g32 = gu.float().requires_grad_(True)
    r = torch.nn.functional.silu(g32[:, :hp]) * g32[:, hp:]
    assert torch.allclose(act.float(), r, rtol=1e-2, atol=1e-2)
    da = torch.randn(M, hp, generator=g).to(torch.bfloat16).to(DEV)
    r.backward(da.float())
    assert torch.allclose(ops.swiglu_bwd(gu, da, hp).float(), g32.grad, rtol=2e-2, atol=2e-2)
    # adamw vs torch.optim.AdamW
    p = torch.randn(1000, generator=g).to(DEV)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    sh = torch.empty(1000, dtype=torch.bfloat16, device=DEV)
    for step in (1, 2, 3):
        gr = torch.randn(1000, generator=g).to(DEV)
        pr.grad = gr.clone()
        opt.step()
        ops.adamw(p, gr, m, v, sh, 3e-4, 0.9, 0.95, 1e-8, 0.05, step)
    assert torch.allclose(p, pr.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(sh, p.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------ heads / CE
def test_skinny_linear_and_ce(ops):
    g = torch.Generator().manual_seed(6)
    for (B, T, d, V) in [(2, 33, 32, 68), (4, 256, 512, 68), (3, 100, 384, 69), (2, 64, 64, 5), (9, 333, 512, 5), (2, 50, 384, 8), (8, 300, 64, 128), (2, 70, 64, 64)]:
        M = B * T
        x = torch.randn(M, d, generator=g).to(DEV)
        w = (torch.randn(V, d, generator=g) * 0.2).to(DEV)
        b = torch.randn(V, generator=g).to(DEV) if V == 5 else None
        out = ops.skinny_linear_fwd(x, w, b)
        ref = x @ w.t() + (b if b is not None else 0)
        assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4)
        dout = torch.randn(M, V, generator=g).to(DEV)
        dx = torch.ones(M, d, device=DEV)
        dw = torch.zeros_like(w)
        db = torch.zeros(V, device=DEV) if b is not None else None
        ops.skinny_linear_bwd(dout, x, w, dx, True, dw, db)
        assert torch.allclose(dx, 1 + dout @ w, rtol=1e-4, atol=1e-4)
        assert torch.allclose(dw, dout.t() @ x, rtol=1e-4, atol=1e-3)
        if db is not None:
            assert torch.allclose(db, dout.sum(0), rtol=1e-4, atol=1e-3)
        # cross entropy vs torch, with ignore / smoothing / class weights
        tgt = torch.randint(0, V, (B, T), generator=g).to(DEV)
        cw = (torch.rand(V, generator=g) + 0.5).to(DEV)
        for eps, cwt in ((0.0, None), (0.05, None), (0.1, cw)):
            lg = (out * 3).contiguous()
            sums, lse, mean = ops.ce_fwd(lg, tgt, B, T, class_w=cwt, smoothing=eps, ignore_index=0)
            assert mean.item() == (sums[0] / sums[1]).item()
            l32 = lg.clone().requires_grad_(True)
            rl = torch.nn.functional.cross_entropy(l32, tgt.view(-1), ignore_index=0, label_smoothing=eps, weight=cwt)
            assert (sums[0] / sums[1]).item() == pytest.approx(rl.item(), rel=2e-6)
            rl.backward()
            dl = ops.ce_bwd(lg, lse, tgt, sums, None, B, T, class_w=cwt, smoothing=eps, ignore_index=0)
            assert torch.allclose(dl, l32.grad, rtol=1e-4, atol=1e-7)
            # bf16 by-products of the same kernel: padded copy, and the hi|lo|hi split (== the stand-alone kernels)
            Vp = (V + 7) // 8 * 8
            dl1, cp = ops.ce_bwd(lg, lse, tgt, sums, None, B, T, class_w=cwt, smoothing=eps, ignore_index=0, bf16_mode=1)
            assert torch.equal(dl1, dl) and torch.equal(cp, ops.cast_bf16(dl, ld_out=Vp))
            dl2, sp = ops.ce_bwd(lg, lse, tgt, sums, None, B, T, class_w=cwt, smoothing=eps, ignore_index=0, bf16_mode=2)
            assert torch.equal(dl2, dl) and torch.equal(sp, ops.split3(dl, cols_pad=Vp))


def test_ce_offset_mask_matches_oracle(ops):
    g = torch.Generator().manual_seed(7)
    B, T, V = 4, 96, 68
    idx, tgt = O.synthetic_batch(B, T, seed=3, realistic=True)
    tgt = tgt.to(DEV)
    logits = torch.randn(B, T, V, generator=g).to(DEV)
    nb = ops.next_in_set(tgt, (2, 3))
    for o in (2, 4, 8, 16):
        sums, _, mean = ops.ce_fwd(logits.view(-1, V), tgt, B, T, shift=o - 1, next_boundary=nb, smoothing=0.05,
                                   zero_if_empty=True)
        total, losses = O.multi_offset_lm_loss({o: logits.cpu()}, tgt.cpu(), {o: 1.0}, label_smoothing=0.05)
        if o in losses:
            assert (sums[0] / sums[1]).item() == pytest.approx(losses[o].item(), rel=3e-6)
        else:
            assert sums[1].item() == 0.0 and mean.item() == 0.0


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, idx, B, T, H, Hk, hd, window, sep_id):
    W = (H + 2 * Hk) * hd
    x = qkv.float().view(B, T, W)
    q = x[..., : H * hd].view(B, T, H, hd).transpose(1, 2)
    k = x[..., H * hd:(H + Hk) * hd].view(B, T, Hk, hd).transpose(1, 2).repeat_interleave(H // Hk, dim=1)
    v = x[..., (H + Hk) * hd:].view(B, T, Hk, hd).transpose(1, 2).repeat_interleave(H // Hk, dim=1)
    m = O.attention_mask(idx.cpu().numpy(), sep_id, window)
    if m is None:
        m = np.tril(np.ones((T, T), dtype=bool))[None, None]
    m = torch.from_numpy(np.ascontiguousarray(m)).to(qkv.device)
    att = (q @ k.transpose(-2, -1)) / math.sqrt(hd)
    att = att.masked_fill(~m, float("-inf"))
    lse = torch.logsumexp(att, dim=-1)
    p = torch.softmax(att, dim=-1)
    y = (p @ v).transpose(1, 2).reshape(B * T, H * hd)
    return y, lse, p


ATTN_CASES = [
    # B, T, H, Hk, hd, window, sep
    (1, 128, 1, 1, 64, None, None),
    (2, 256, 2, 2, 64, None, 3),
    (2, 200, 4, 2, 64, None, 3),
    (1, 384, 2, 1, 32, None, 3),
    (2, 130, 8, 4, 48, None, 3),
    (1, 256, 2, 2, 128, None, 3),
    (2, 300, 2, 2, 64, 37, 3),
    (1, 40, 1, 1, 16, None, 3),
    (1, 1024, 8, 8, 64, None, 3),
    (3, 333, 2, 2, 64, None, 3),   # B*H*T not a multiple of 4: workspace alignment
    (1, 1, 1, 1, 32, None, 3),
    (2, 129, 2, 1, 64, 1, None),   # window 1 = identity attention
    # C5 (BASELINE configs[4]): long-context causal attention, seq 4096, 8 heads, hd 48 and 64, segment-causal / causal
    (1, 4096, 8, 8, 48, None, 3),
    (1, 4096, 8, 8, 64, None, None),
    (1, 4096, 8, 4, 64, 1000, 3),
]


@pytest.mark.parametrize("case", ATTN_CASES, ids=[str(c) for c in ATTN_CASES])
def test_attention_fwd_bwd(ops, case):
    B, T, H, Hk, hd, window, sep = case
    g = torch.Generator().manual_seed(8)
    idx, _ = O.synthetic_batch(B, T, seed=5, realistic=True)
    idx = idx.to(DEV)
    qkv = (torch.randn(B * T, (H + 2 * Hk) * hd, generator=g) * 1.0).to(torch.bfloat16).to(DEV)
    ss = ops.segment_starts(idx, sep) if sep is not None else None
    out, lse = ops.attn_fwd(qkv, ss, B, T, H, Hk, hd, window=window or 0)
    torch.cuda.synchronize()
    q32 = qkv.float().requires_grad_(True)
    ry, rlse, rp = _attn_ref(q32, idx, B, T, H, Hk, hd, window, sep)
    err = (out.float() - ry).abs().max().item()
    assert err <= 2e-2, f"fwd max-abs err {err}"
    assert torch.allclose(lse, rlse, rtol=1e-3, atol=2e-3)
    probs = ops.attn_probs(qkv, ss, B, T, H, Hk, hd, window=window or 0)
    assert torch.allclose(probs, rp, rtol=1e-3, atol=1e-4)
    dout = (torch.randn(B * T, H * hd, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    ry.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, ss, out, dout, lse, B, T, H, Hk, hd, window=window or 0)
    torch.cuda.synchronize()
    W = (H + 2 * Hk) * hd
    for name, sl in (("dq", slice(0, H * hd)), ("dk", slice(H * hd, (H + Hk) * hd)), ("dv", slice((H + Hk) * hd, W))):
        a, r = dqkv.float()[:, sl], q32.grad[:, sl]
        # identity attention (T=1 or window=1) has dq = dk = 0 exactly: gate on the dO scale there
        rel = ((a - r).norm() / max(r.norm().item(), 1e-3 * dout.float().norm().item())).item()
        assert rel <= 2e-2, f"{name} rel-norm err {rel}"


@pytest.mark.parametrize("hd", [64, 32])
def test_attention_forward_reference_maximum_is_moved_when_scores_grow(ops, hd):
    """The forward keeps the row maximum of the FIRST visible kv tile as the softmax reference and moves it only when a
    later tile overshoots it by ~2^40 (attn_fwd_w3_kernel).  Scores that grow by hundreds of nats along the key axis
    force that path on most tiles; scores that shrink exercise the opposite side (late tiles underflow to exact
    zeros).  Output and LSE must match the fp32 reference either way, and the backward must accept that LSE."""
    B, T, H = 2, 640, 2
    g = torch.Generator().manual_seed(3)
    u = torch.nn.functional.normalize(torch.randn(hd, generator=g), dim=0)
    pos = torch.arange(T, dtype=torch.float32) / T
    for direction in (+1.0, -1.0):
        q = (u * 8.0)[None, :].repeat(T, 1) + 0.05 * torch.randn(T, hd, generator=g)
        k = (direction * 300.0 * pos)[:, None] * u[None, :] + 0.05 * torch.randn(T, hd, generator=g)
        v = torch.randn(T, hd, generator=g)
        row = torch.cat([q.repeat(1, H), k.repeat(1, H), v.repeat(1, H)], dim=1)
        qkv = row.repeat(B, 1).to(torch.bfloat16).to(DEV)
        idx = torch.full((B, T), 5, dtype=torch.long, device=DEV)
        out, lse = ops.attn_fwd(qkv, None, B, T, H, H, hd)
        q32 = qkv.float().requires_grad_(True)
        ry, rlse, _ = _attn_ref(q32, idx, B, T, H, H, hd, None, None)
        assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
        assert (out.float() - ry).abs().max().item() <= 3e-2
        assert torch.allclose(lse, rlse, rtol=1e-3, atol=5e-3)
        dout = torch.randn(B * T, H * hd, generator=g).to(torch.bfloat16).to(DEV)
        ry.backward(dout.float())
        dqkv = ops.attn_bwd(qkv, None, out, dout, lse, B, T, H, H, hd)
        dv = dqkv.float()[:, 2 * H * hd:]
        rdv = q32.grad[:, 2 * H * hd:]
        assert ((dv - rdv).norm() / rdv.norm()).item() <= 2e-2


def test_split_head_on_tensor_cores_is_fp32_accurate(ops):
    """hi/lo bf16 split head (tcgen05) vs float64: forward, dx, dW, db — errors must sit at the 1e-5 level,
    two orders below a plain bf16 GEMM."""
    from codonlm_b200 import functional as Fn
    g = torch.Generator().manual_seed(11)
    for (M, d, V, has_b) in [(4096, 512, 68, False), (4100, 128, 8, True)]:
        x = torch.randn(M, d, generator=g).to(DEV).requires_grad_(True)
        w = (torch.randn(V, d, generator=g) * 0.05).to(DEV).requires_grad_(True)
        b = torch.randn(V, generator=g).to(DEV).requires_grad_(True) if has_b else None
        out = Fn.SplitHeadFn.apply(x, w, b)
        ref = x.double() @ w.double().t() + (b.double() if has_b else 0)
        scale = ref.abs().max().item()
        assert (out.double() - ref).abs().max().item() <= 3e-5 * scale
        go = torch.randn(M, V, generator=g).to(DEV)
        out.backward(go)
        rdx = go.double() @ w.double()
        rdw = go.double().t() @ x.double()
        assert (x.grad.double() - rdx).abs().max().item() <= 3e-5 * rdx.abs().max().item()
        assert (w.grad.double() - rdw).abs().max().item() <= 3e-5 * rdw.abs().max().item()
        if has_b:
            assert (b.grad.double() - go.double().sum(0)).abs().max().item() <= 3e-5 * go.double().sum(0).abs().max().item()


@pytest.mark.parametrize("case", [(2, 300, 4, 4, 64, 0), (1, 257, 4, 2, 32, 0), (2, 130, 2, 1, 48, 40), (1, 200, 2, 2, 128, 0)])
def test_attention_backward_fused_bias_column_sums(ops, case):
    """cgpt_attn_bwd_colsum: the q|k|v bias gradients taken inside the backward kernels equal the column sums of
    the dqkv they wrote (same bf16-rounded values; fp32 accumulation order differs), and accumulate into the buffer."""
    B, T, H, Hk, hd, window = case
    g = torch.Generator().manual_seed(5)
    W = (H + 2 * Hk) * hd
    qkv = torch.randn(B * T, W, generator=g).to(DEV).to(torch.bfloat16)
    idx = torch.randint(4, 68, (B, T), generator=g)
    idx[:, T // 3] = 3
    ss = ops.segment_starts(idx.to(DEV), 3)
    out, lse = ops.attn_fwd(qkv, ss, B, T, H, Hk, hd, window=window)
    dout = torch.randn(B * T, H * hd, generator=g).to(DEV).to(torch.bfloat16)
    base = torch.full((W,), 0.5, device=DEV)
    csum = base.clone()
    dqkv = ops.attn_bwd(qkv, ss, out, dout, lse, B, T, H, Hk, hd, window=window, colsum=csum)
    ref = dqkv.float().sum(0)
    assert torch.equal(dqkv, ops.attn_bwd(qkv, ss, out, dout, lse, B, T, H, Hk, hd, window=window))
    err = (csum - base - ref).abs().max().item()
    assert err <= 1e-4 * max(1.0, ref.abs().max().item()), err


# ------------------------------------------------------------------------------------------ dropout
def test_elementwise_dropout_mask_statistics_and_backward(ops):
    n = 1 << 20
    x = torch.ones(n, device=DEV)
    res = torch.full((n,), 2.0, device=DEV)
    for p in (0.1, 0.5):
        y = ops.dropout(x, None, p, seed=1234, offset=8)
        keep = (y != 0).float().mean().item()
        assert abs(keep - (1 - p)) < 4e-3
        assert torch.allclose(y[y != 0], torch.full_like(y[y != 0], 1 / (1 - p)))
        assert torch.equal(y, ops.dropout(x, None, p, seed=1234, offset=8))          # pure function of (seed, offset)
        assert not torch.equal(y, ops.dropout(x, None, p, seed=1234, offset=12))
        assert not torch.equal(y, ops.dropout(x, None, p, seed=1235, offset=8))
        yr = ops.dropout(x, res, p, seed=1234, offset=8)
        assert torch.equal(yr, y + 2.0)
        g = torch.randn(n, device=DEV)
        gb = ops.dropout(g, None, p, seed=1234, offset=8, out_bf16=True)             # backward: same mask on the grad
        assert torch.equal(gb, (g * y).to(torch.bfloat16))


@pytest.mark.parametrize("case", [(2, 256, 2, 2, 64, None, 3), (1, 200, 4, 2, 32, None, 3), (2, 130, 2, 1, 48, 40, 3)])
def test_attention_dropout_matches_masked_reference(ops, case):
    """Dropout on the attention probabilities: extract the kernel's own Philox mask through the dense
    probability kernel, then compare fwd and bwd with an fp32 reference that uses that mask."""
    B, T, H, Hk, hd, window, sep = case
    p, seed, off = 0.2, 77, 16
    g = torch.Generator().manual_seed(21)
    idx, _ = O.synthetic_batch(B, T, seed=6, realistic=True)
    idx = idx.to(DEV)
    qkv = torch.randn(B * T, (H + 2 * Hk) * hd, generator=g).to(torch.bfloat16).to(DEV)
    ss = ops.segment_starts(idx, sep)
    kw = dict(window=window or 0)
    pr0 = ops.attn_probs(qkv, ss, B, T, H, Hk, hd, **kw)
    pr1 = ops.attn_probs(qkv, ss, B, T, H, Hk, hd, dropout_p=p, seed=seed, offset=off, **kw)
    vis = pr0 > 0
    mask = torch.where(vis, pr1 / pr0.clamp_min(1e-30), torch.zeros_like(pr0))       # 0 or 1/(1-p)
    kept = (mask[vis] > 0).float().mean().item()
    assert abs(kept - (1 - p)) < 2e-2
    # attention masks draw 16 bits per element: the effective drop probability is floor(p 2^32) >> 16 over 2^16, and
    # the kept probabilities are scaled by exactly 1 / (1 - that), so the mask stays unbiased
    t16 = int(p * 2 ** 32) >> 16
    assert torch.allclose(mask[vis & (mask > 0)], torch.tensor(65536.0 / (65536 - t16), device=DEV), rtol=1e-5)
    out, lse = ops.attn_fwd(qkv, ss, B, T, H, Hk, hd, dropout_p=p, seed=seed, offset=off, **kw)
    out2, _ = ops.attn_fwd(qkv, ss, B, T, H, Hk, hd, dropout_p=p, seed=seed, offset=off, **kw)
    assert torch.equal(out, out2)
    q32 = qkv.float().requires_grad_(True)
    _, _, probs = _attn_ref(q32, idx, B, T, H, Hk, hd, window, sep)
    W = (H + 2 * Hk) * hd
    v = q32.view(B, T, W)[..., (H + Hk) * hd:].view(B, T, Hk, hd).transpose(1, 2).repeat_interleave(H // Hk, dim=1)
    ry = ((probs * mask) @ v).transpose(1, 2).reshape(B * T, H * hd)
    assert (out.float() - ry).abs().max().item() <= 3e-2
    dout = (torch.randn(B * T, H * hd, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    ry.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, ss, out, dout, lse, B, T, H, Hk, hd, dropout_p=p, seed=seed, offset=off, **kw)
    for name, sl in (("dq", slice(0, H * hd)), ("dk", slice(H * hd, (H + Hk) * hd)), ("dv", slice((H + Hk) * hd, W))):
        a, r = dqkv.float()[:, sl], q32.grad[:, sl]
        rel = ((a - r).norm() / r.norm()).item()
        assert rel <= 3e-2, f"{name} rel-norm err {rel}"


def test_cross_entropy_out_of_range_label_poisons_the_loss(ops):
    """A label outside [0, V) that is not ignore_index: F.cross_entropy raises; the kernels must not read out of bounds
    and must not return a finite loss — the loss is NaN (the trainer's non-finite policy takes it from there), the
    gradient of that row is zero, every other row is unaffected."""
    from codonlm_b200 import functional as Fn
    g = torch.Generator().manual_seed(2)
    B, T, V = 2, 16, 68
    logits = torch.randn(B * T, V, generator=g).to(DEV).requires_grad_(True)
    tgt = torch.randint(1, V, (B, T), generator=g).to(DEV)
    good, _ = Fn.CrossEntropyFn.apply(logits, tgt, None, None, B, T, 0, 0.05, 0, False)
    assert torch.isfinite(good)
    bad_t = tgt.clone()
    bad_t[1, 3] = V + 5
    bad_t[0, 7] = -7
    loss, _ = Fn.CrossEntropyFn.apply(logits, bad_t, None, None, B, T, 0, 0.05, 0, False)
    assert torch.isnan(loss)
