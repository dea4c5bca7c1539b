"""Autograd glue: each Function's forward/backward is a short sequence of C-ABI kernel launches.

Activations between Functions are 2-D [M=B*T, features]; the residual stream is fp32, GEMM operands
are bf16, weight gradients are fp32.  No arithmetic is done by PyTorch here except trivial scalar
bookkeeping (loss = sum / count) — torch provides memory, streams and the autograd graph.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import ops
from .ops import EPI_GELU, EPI_MUL_AUX, EPI_NONE, bf16, f32

_SMS = 148

# By-products of a backward kernel that the NEXT backward node wants (LayerNorm backward -> the residual linear in
# front of it: the bf16 copy and the column sums of dx; attention backward -> the QKV linear: the column sums of dqkv)
# travel ON the gradient tensor itself, as a Python attribute: autograd hands the same tensor object to the next node,
# and when it does not (a hook, an accumulation, a clone in between) the attribute is simply gone and the consumer
# recomputes.  Nothing is keyed by device address, nothing is module-global: two models, retain_graph or a second
# backward in flight cannot cross wires.
_SIDE_ATTR = "_cgpt_side"


def _attach(g: torch.Tensor, bf16_copy, colsum):
    setattr(g, _SIDE_ATTR, (bf16_copy, colsum))


def _detach_side(g: torch.Tensor):
    """By-products attached to `g` — or to the tensor `g` is a full, contiguous view of: between two blocks the residual
    stream passes reshape nodes ((B,T,d) <-> (B*T,d)), whose backward hands on a VIEW of the producer's tensor."""
    hit = getattr(g, _SIDE_ATTR, None)
    if hit is not None:
        delattr(g, _SIDE_ATTR)
        return hit
    base = g._base
    if base is not None and base.data_ptr() == g.data_ptr() and base.numel() == g.numel() and g.is_contiguous():
        hit = getattr(base, _SIDE_ATTR, None)
        if hit is not None:
            delattr(base, _SIDE_ATTR)
    return hit


def _grad_form_of(g: torch.Tensor, shape):
    """bf16 form of the logit gradient `g` left by the cross-entropy backward (padded copy or hi|lo|hi split), if it
    travelled with `g` and has the wanted shape."""
    hit = _detach_side(g)
    if hit is not None and hit[0] is not None and tuple(hit[0].shape) == tuple(shape):
        return hit[0]
    return None


def _bf16_of(g: torch.Tensor):
    """(bf16 copy of g, column sums of g or None)."""
    hit = _detach_side(g)
    if hit is not None and hit[0] is not None and hit[0].shape == g.shape:
        return hit
    return ops.cast_bf16(g), None


# When the producer kernel has ALREADY added the column sums to `param.main_grad` (no scratch vector, no add kernel)
# it leaves a credit on the Parameter object; the consumer that owns that bias takes the credit instead of computing
# the sums.  Credits and debits balance within every backward pass (the producer always runs before its consumer);
# trainer.TrainStep.zero_grad() clears leftovers of an interrupted backward.
def _credit(param):
    param._cgpt_direct = getattr(param, "_cgpt_direct", 0) + 1


def _take_credit(param) -> bool:
    n = getattr(param, "_cgpt_direct", 0) if param is not None else 0
    if n > 0:
        param._cgpt_direct = n - 1
        return True
    return False


def clear_credits(params):
    for p in params:
        if getattr(p, "_cgpt_direct", 0):
            p._cgpt_direct = 0


def _bias_grad(gb: torch.Tensor, n: int, master, colsum, off: int = 0):
    """Bias gradient = column sums of the output gradient: taken from the producer when it already has them."""
    if _take_credit(master):  # the producer kernel added them to master.main_grad itself
        _done(master)
        return None
    if colsum is None:
        return _colsum(gb, n, master=master, off=off)
    mg = _main_grad(master)
    if mg is not None:
        mg.add_(colsum)
        _done(master)
        return None
    return colsum


def _main_grad(p):
    """Flat-buffer gradient slot installed by trainer.TrainStep (None when the model is used with a plain
    torch optimiser): kernels accumulate into it directly and autograd gets no gradient tensor."""
    return getattr(p, "main_grad", None) if p is not None else None


def _done(p):
    cb = getattr(p, "_cgpt_grad_ready", None)
    if cb is not None:
        cb()


def _wgrad(dy: torch.Tensor, x: torch.Tensor, n_out: int, k_in: int, dy_ld=None, x_ld=None, dy_off=0, master=None):
    """dW[n_out, k_in] (+)= dy[:, dy_off:dy_off+n_out]ᵀ · x[:, :k_in]  (fp32, split-K over the tokens).
    Accumulates into master.main_grad when present (returns None), else returns a fresh tensor."""
    Mtok = x.shape[0]
    mg = _main_grad(master)
    dw = mg if mg is not None else torch.zeros((n_out, k_in), dtype=f32, device=x.device)
    a = dy if dy_off == 0 else dy[:, dy_off:]
    tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
    split = ops.pick_split_k(tiles, (Mtok + 63) // 64, _SMS)
    ops.gemm(a, x, dw, M=n_out, N=k_in, K=Mtok, a_mn=True, b_mn=True, lda=dy_ld or dy.stride(0),
             ldb=x_ld or x.stride(0), ldc=k_in, accumulate=True, split_k=split)
    if mg is not None:
        _done(master)
        return None
    return dw


def _wgrad_into(dy: torch.Tensor, x: torch.Tensor, n_out: int, k_in: int, dw: torch.Tensor):
    """dw[n_out, k_in] += dyᵀ · x into an existing fp32 matrix (split-K over the tokens)."""
    Mtok = x.shape[0]
    tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
    split = ops.pick_split_k(tiles, (Mtok + 63) // 64, _SMS)
    ops.gemm(dy, x, dw, M=n_out, N=k_in, K=Mtok, a_mn=True, b_mn=True, lda=dy.stride(0), ldb=x.stride(0), ldc=k_in,
             accumulate=True, split_k=split)


def _packed_slots(masters, rows):
    """main_grad slots of `masters` when all exist, lie back to back in memory and cover `rows` without gaps."""
    slots = [_main_grad(m) for m in masters]
    if not slots or any(s is None for s in slots):
        return None
    r = 0
    for s, (r0, n) in zip(slots, rows):
        if r0 != r or s.shape[0] != n:
            return None
        r += n
    for a, b in zip(slots[:-1], slots[1:]):
        if a.data_ptr() + a.numel() * a.element_size() != b.data_ptr():
            return None
    return slots


def _adjacent_tensors(tensors) -> bool:
    """True when the tensors sit back to back in memory (same dtype), i.e. form one packed row-major matrix."""
    for a, b in zip(tensors[:-1], tensors[1:]):
        if a.dtype != b.dtype or a.data_ptr() + a.numel() * a.element_size() != b.data_ptr():
            return False
    return True


def _colsum(dy: torch.Tensor, n: int, master=None, off: int = 0):
    mg = _main_grad(master)
    out = mg if mg is not None else torch.zeros((n,), dtype=f32, device=dy.device)
    ops.colsum_bf16(dy if off == 0 else dy[:, off:], out, N=n, ld=dy.stride(0))
    if mg is not None:
        _done(master)
        return None
    return out


class EmbedFn(Function):
    """x = tok_emb[idx] (+ pos_emb[:T])  — model_tiny_gpt.py:306-309."""

    @staticmethod
    def forward(ctx, idx, tok_w, pos_w):
        ctx.save_for_backward(idx)
        ctx.masters = (tok_w, pos_w)
        return ops.embed_fwd(idx, tok_w, pos_w)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        tok_w, pos_w = ctx.masters
        g = g.contiguous()
        mt, mp = _main_grad(tok_w), _main_grad(pos_w)
        dtok = mt if mt is not None else torch.zeros(tok_w.shape, dtype=f32, device=g.device)
        dpos = None
        if pos_w is not None:
            dpos = mp if mp is not None else torch.zeros(pos_w.shape, dtype=f32, device=g.device)
        ops.embed_bwd(idx, g, dtok, dpos)
        if mt is not None:
            _done(tok_w)
        if mp is not None:
            _done(pos_w)
        return None, (None if mt is not None else dtok), (None if mp is not None else dpos)


class ShapeProjFn(Function):
    """x + shape_proj(shape_embeddings) — nn.Linear(3, d) on the DNA-shape features added to the embedding
    (model_tiny_gpt.py:226-229, 310-311).  K = 3 is no tensor-core shape: three small memory-bound kernels."""

    @staticmethod
    def forward(ctx, x, s, w, b):
        ctx.save_for_backward(s, w)
        ctx.masters = (w, b)
        return ops.shape_proj_fwd(x, s, w.detach(), b.detach())

    @staticmethod
    def backward(ctx, g):
        s, w = ctx.saved_tensors
        wm, bm = ctx.masters
        g = g.contiguous()
        mw, mb = _main_grad(wm), _main_grad(bm)
        dw = mw if mw is not None else torch.zeros_like(w)
        db = mb if mb is not None else torch.zeros((w.shape[0],), dtype=f32, device=w.device)
        ds = ops.shape_proj_bwd(g, s, w.detach(), dw, db, ctx.needs_input_grad[1])
        if mw is not None:
            _done(wm)
            _done(bm)
            return g, ds, None, None
        return g, ds, dw, db


class ResidualLayerNormFn(Function):
    """(x, y) = (x, LN(x)): returning the residual stream through the Function lets backward fuse
    dx = d_residual + LN'(dy) into one kernel (model_tiny_gpt.py:151-152 pre-norm pattern)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, want_f32, colsum_target=None):
        """colsum_target: the bias Parameter of the residual linear right in front of this LayerNorm (its gradient
        is the column sum of this node's dx), or None.  With a flat gradient buffer the backward kernel adds straight
        into its slot."""
        M, d = x.shape
        # outputs nobody differentiates through (the residual alias and the bf16 copy of ln_f, whose fp32 output feeds
        # the heads) must reach backward as None, not as zero tensors autograd would have to fill, convert and add
        ctx.set_materialize_grads(False)
        yb, yf, mean, rstd = ops.layernorm_fwd(x, gamma, beta, want_bf16=True, want_f32=want_f32)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.masters = (gamma, beta)
        ctx.want_f32 = want_f32
        ctx.colsum_target = colsum_target
        if want_f32:
            return x.view_as(x), yb, yf
        return x.view_as(x), yb

    @staticmethod
    def backward(ctx, gx, gyb, gyf=None):
        x, gamma, mean, rstd = ctx.saved_tensors
        gam_p, bet_p = ctx.masters
        mg, mb = _main_grad(gam_p), _main_grad(bet_p)
        dgamma = mg if mg is not None else torch.zeros_like(gamma)
        dbeta = mb if mb is not None else torch.zeros_like(gamma)
        if gyf is not None and gyb is not None:
            dy = gyf + gyb.float()
        elif gyf is not None:
            dy = gyf
        elif gyb is not None:
            dy = gyb
        else:
            return gx, None, None, None, None
        tgt = ctx.colsum_target
        tmg = _main_grad(tgt)
        direct = tmg is not None and tmg.numel() == gamma.numel() and tmg.is_contiguous()
        dxsum = tmg if direct else torch.zeros_like(gamma)  # direct: the kernel adds into the bias' gradient slot itself
        dx, dxb = ops.layernorm_bwd(dy.contiguous(), x, gamma, mean, rstd, None if gx is None else gx.contiguous(),
                                    dgamma, dbeta, want_bf16=True, dx_colsum=dxsum)
        # the upstream residual GEMM's backward wants dx as a bf16 operand and its column sums as the bias gradient
        if direct:
            _credit(tgt)
            _attach(dx, dxb, None)
        else:
            _attach(dx, dxb, dxsum)
        if mg is not None:
            _done(gam_p)
            _done(bet_p)
            return dx, None, None, None, None
        return dx, dgamma, dbeta, None, None


class PackedLinearFn(Function):
    """y = x·W_packedᵀ + b_packed (+ residual): one GEMM for several nn.Linear that share the input
    (query|key|value, model_tiny_gpt.py:85-93; w_gate|w_up, :57) or for a single one (proj :132).

    `masters` are the fp32 Parameters in packed row order: weights first, then biases (may be empty);
    `rows` lists (row_offset_in_packed, n_rows) per weight."""

    @staticmethod
    def forward(ctx, x, w_sh, b_sh, residual, rows, out_f32, *masters):
        M, K = x.shape
        N = w_sh.shape[0]
        out = torch.empty((M, N), dtype=f32 if (out_f32 or residual is not None) else bf16, device=x.device)
        ops.gemm(x, w_sh, out, M=M, N=N, K=K, bias=b_sh, residual=residual)
        ctx.save_for_backward(x, w_sh)
        ctx.rows = rows
        ctx.has_res = residual is not None
        ctx.n_w = len(rows)
        ctx.has_bias = b_sh is not None
        ctx.k_in = [m.shape[1] for m in masters[: len(rows)]]
        ctx.masters = masters
        return out

    @staticmethod
    def backward(ctx, g):
        x, w_sh = ctx.saved_tensors
        M, K = x.shape
        N = w_sh.shape[0]
        g = g.contiguous()
        if g.dtype == f32:
            gb, gsum = _bf16_of(g)
        else:  # bf16 gradient (dqkv from the attention backward): its column sums may ride along
            hit = _detach_side(g)
            gb, gsum = g, (None if hit is None else hit[1])
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=bf16, device=x.device)
            ops.gemm(gb, w_sh, dx, M=M, N=K, K=N, b_mn=True)
        nw = len(ctx.rows)
        grads: List[Optional[torch.Tensor]] = []
        # gradient slots that sit back to back in the flat buffer (trainer._packed_order) form one [N, K] matrix:
        # one wgrad GEMM / one column sum instead of one per nn.Linear
        w_slots = _packed_slots(ctx.masters[:nw], ctx.rows)
        if nw > 1 and w_slots is not None and len(set(ctx.k_in)) == 1:
            packed = torch.as_strided(w_slots[0], (N, K), (K, 1))
            _wgrad_into(gb, x, N, K, packed)
            for m in ctx.masters[:nw]:
                _done(m)
            grads.extend([None] * nw)
        else:
            for i, ((r0, n), k_in) in enumerate(zip(ctx.rows, ctx.k_in)):
                grads.append(_wgrad(gb, x, n, k_in, dy_off=r0, master=ctx.masters[i]))
        if ctx.has_bias:
            b_slots = _packed_slots(ctx.masters[nw:], ctx.rows)
            if nw > 1 and b_slots is not None:
                packed_b = torch.as_strided(b_slots[0], (N,), (1,))
                if _take_credit(ctx.masters[nw]):
                    pass  # the attention backward kernels already added the column sums of dqkv to these slots
                elif gsum is not None and gsum.numel() == N:
                    packed_b.add_(gsum)
                else:
                    ops.colsum_bf16(gb, packed_b, N=N, ld=gb.stride(0))
                for m in ctx.masters[nw:]:
                    _done(m)
                grads.extend([None] * nw)
            else:
                if gsum is not None and gsum.numel() != N:
                    gsum = None
                for i, (r0, n) in enumerate(ctx.rows):
                    grads.append(_bias_grad(gb, n, ctx.masters[nw + i], None if gsum is None else gsum[r0:r0 + n],
                                            off=r0))
        return (dx, None, None, g if ctx.has_res else None, None, None, *grads)


class MlpGeluFn(Function):
    """x + fc2(gelu(fc1(h)))  — model_tiny_gpt.py:143-148,152 with bias+GELU and bias+residual fused into
    the GEMM epilogues; the fc1 epilogue also emits gelu'(pre), which the fc2 dgrad epilogue multiplies in."""

    @staticmethod
    def forward(ctx, h, x_res, w1_sh, b1, w2_sh, b2, w1, w2):
        M, d = h.shape
        F = w1_sh.shape[0]
        # gelu'(pre), written by the same epilogue — only when a backward pass can follow (not under no_grad)
        dact = torch.empty((M, F), dtype=bf16, device=h.device) if any(ctx.needs_input_grad) else None
        act = torch.empty((M, F), dtype=bf16, device=h.device)
        ops.gemm(h, w1_sh, act, M=M, N=F, K=d, bias=b1, epilogue=EPI_GELU, aux_out=dact, ldaux=F)
        n_out = w2_sh.shape[0]
        out = torch.empty((M, n_out), dtype=f32, device=h.device)
        ops.gemm(act, w2_sh, out, M=M, N=n_out, K=F, bias=b2, residual=x_res)
        ctx.save_for_backward(h, dact, act, w1_sh, w2_sh)
        ctx.has_res = x_res is not None
        ctx.masters = (b1, b2, w1, w2)
        return out

    @staticmethod
    def backward(ctx, g):
        h, dact, act, w1_sh, w2_sh = ctx.saved_tensors
        M, d = h.shape
        F = w1_sh.shape[0]
        n_out = w2_sh.shape[0]
        g = g.contiguous()
        b1, b2, w1, w2 = ctx.masters
        gb, gsum = _bf16_of(g)
        dw2 = _wgrad(gb, act, n_out, F, master=w2)
        db2 = _bias_grad(gb, n_out, b2, gsum)
        dpre = torch.empty((M, F), dtype=bf16, device=h.device)
        fuse = (F % 8 == 0)  # TMA-aligned: the GEMM epilogue also produces the fc1 bias gradient
        mb1 = _main_grad(b1)
        db1 = (mb1 if mb1 is not None else torch.zeros((F,), dtype=f32, device=h.device)) if fuse else None
        ops.gemm(gb, w2_sh, dpre, M=M, N=F, K=n_out, b_mn=True, epilogue=EPI_MUL_AUX, aux=dact, ldaux=F, colsum=db1)
        dw1 = _wgrad(dpre, h, F, d, master=w1)
        if not fuse:
            db1 = _colsum(dpre, F, master=b1)
        elif mb1 is not None:
            _done(b1)
            db1 = None
        dh = torch.empty((M, d), dtype=bf16, device=h.device)
        ops.gemm(dpre, w1_sh, dh, M=M, N=d, K=F, b_mn=True)
        return dh, (g if ctx.has_res else None), None, db1, None, db2, dw1, dw2


class MlpSwiGLUFn(Function):
    """x + w_down(silu(w_gate h) * w_up h)  — model_tiny_gpt.py:47-57,152.  Hidden width h=int(8d/3) is
    padded to hp (multiple of 8) in the bf16 shadows; the padding columns are exact zeros."""

    @staticmethod
    def forward(ctx, hin, x_res, wgu_sh, wd_sh, w_gate, w_up, w_down):
        M, d = hin.shape
        hp = wgu_sh.shape[0] // 2
        gu = torch.empty((M, 2 * hp), dtype=bf16, device=hin.device)
        ops.gemm(hin, wgu_sh, gu, M=M, N=2 * hp, K=d)
        act = ops.swiglu_fwd(gu, hp)
        n_out = wd_sh.shape[0]
        out = torch.empty((M, n_out), dtype=f32, device=hin.device)
        ops.gemm(act, wd_sh, out, M=M, N=n_out, K=hp, residual=x_res)
        ctx.save_for_backward(hin, gu, act, wgu_sh, wd_sh)
        ctx.hidden = w_gate.shape[0]
        ctx.has_res = x_res is not None
        ctx.masters = (w_gate, w_up, w_down)
        return out

    @staticmethod
    def backward(ctx, g):
        hin, gu, act, wgu_sh, wd_sh = ctx.saved_tensors
        M, d = hin.shape
        hp = wgu_sh.shape[0] // 2
        hid = ctx.hidden
        n_out = wd_sh.shape[0]
        g = g.contiguous()
        w_gate, w_up, w_down = ctx.masters
        gb, _ = _bf16_of(g)
        dwd = _wgrad(gb, act, n_out, hid, master=w_down)        # [d, hid] (unpadded, odd pitch allowed)
        dact = torch.empty((M, hp), dtype=bf16, device=hin.device)
        ops.gemm(gb, wd_sh, dact, M=M, N=hp, K=n_out, b_mn=True)
        dgu = ops.swiglu_bwd(gu, dact, hp)
        dwg = _wgrad(dgu, hin, hid, d, dy_off=0, master=w_gate)
        dwu = _wgrad(dgu, hin, hid, d, dy_off=hp, master=w_up)
        dh = torch.empty((M, d), dtype=bf16, device=hin.device)
        ops.gemm(dgu, wgu_sh, dh, M=M, N=d, K=2 * hp, b_mn=True)
        return dh, (g if ctx.has_res else None), None, None, dwg, dwu, dwd


def _offset_mlp_fwd(xb, w1_sh, b1, w2_sh, b2, need_bwd=True, out_dtype=f32):
    M, d = xb.shape
    dact = torch.empty((M, d), dtype=bf16, device=xb.device) if need_bwd else None
    act = torch.empty((M, d), dtype=bf16, device=xb.device)
    ops.gemm(xb, w1_sh, act, M=M, N=d, K=d, bias=b1, epilogue=EPI_GELU, aux_out=dact, ldaux=d)
    out = torch.empty((M, d), dtype=out_dtype, device=xb.device)
    ops.gemm(act, w2_sh, out, M=M, N=d, K=d, bias=b2)
    return out, act, dact


def _offset_mlp_bwd(gb, xb, dact, act, w1_sh, w2_sh, b1, b2, w1, w2, dx_into=None, act_ld=None, dact_ld=None):
    """gb: bf16 gradient of the MLP output -> (dx bf16, db1, db2, dw1, dw2), None where accumulated in place.
    With `dx_into` (fp32 [M, d]) the input gradient is ADDED to it by the GEMM instead (returns dx None).
    act / dact may be column slices of a wider buffer (row pitch act_ld / dact_ld)."""
    M, d = xb.shape
    dw2 = _wgrad(gb, act, d, d, master=w2, x_ld=act_ld)
    db2 = _colsum(gb, d, master=b2)
    dpre = torch.empty((M, d), dtype=bf16, device=xb.device)
    fuse = (d % 8 == 0)
    mb1 = _main_grad(b1)
    db1 = (mb1 if mb1 is not None else torch.zeros((d,), dtype=f32, device=xb.device)) if fuse else None
    ops.gemm(gb, w2_sh, dpre, M=M, N=d, K=d, b_mn=True, epilogue=EPI_MUL_AUX, aux=dact, ldaux=dact_ld or d, colsum=db1)
    dw1 = _wgrad(dpre, xb, d, d, master=w1)
    if not fuse:
        db1 = _colsum(dpre, d, master=b1)
    elif mb1 is not None:
        _done(b1)
        db1 = None
    if dx_into is not None:
        ops.gemm(dpre, w1_sh, dx_into, M=M, N=d, K=d, b_mn=True, accumulate=True)
        return None, db1, db2, dw1, dw2
    dx = torch.empty((M, d), dtype=bf16, device=xb.device)
    ops.gemm(dpre, w1_sh, dx, M=M, N=d, K=d, b_mn=True)
    return dx, db1, db2, dw1, dw2


class OffsetHeadFn(Function):
    """h_o = W2·gelu(W1·x + b1) + b2 on the post-ln_f hidden state — model_tiny_gpt.py:235-239,335.
    Input bf16, output fp32 (it feeds the fp32 LM head)."""

    @staticmethod
    def forward(ctx, xb, w1_sh, b1, w2_sh, b2, w1, w2):
        out, act, dact = _offset_mlp_fwd(xb, w1_sh, b1, w2_sh, b2, any(ctx.needs_input_grad))
        ctx.save_for_backward(xb, dact, act, w1_sh, w2_sh)
        ctx.masters = (b1, b2, w1, w2)
        return out

    @staticmethod
    def backward(ctx, g):
        xb, dact, act, w1_sh, w2_sh = ctx.saved_tensors
        b1, b2, w1, w2 = ctx.masters
        gb = ops.cast_bf16(g.contiguous())
        dx, db1, db2, dw1, dw2 = _offset_mlp_bwd(gb, xb, dact, act, w1_sh, w2_sh, b1, b2, w1, w2)
        return dx, None, db1, None, db2, dw1, dw2


def reset_side_channel():
    _HEAD_W_CACHE.clear()


def reset_head_cache():
    """Drop the split copies of the head weights (call after writing a master behind autograd's back)."""
    _HEAD_W_CACHE.clear()


class DropoutFn(Function):
    """y = (residual) + dropout(x): nn.Dropout on the embedding (model_tiny_gpt.py:312) and on the MLP branch
    (:57,147) with the residual add fused.  The Philox mask is a function of (seed, offset) drawn from torch's
    CUDA generator; backward re-applies it to the gradient, nothing is stored."""

    @staticmethod
    def forward(ctx, x, residual, p):
        seed, off = ops.philox_state(x.device, x.numel())
        ctx.rng = (seed, off, p)
        ctx.has_res = residual is not None
        return ops.dropout(x.contiguous(), None if residual is None else residual.contiguous(), p, seed, off)

    @staticmethod
    def backward(ctx, g):
        seed, off, p = ctx.rng
        g = g.contiguous()
        _detach_side(g)  # by-products describe g, not dropout(g); the residual path has no consumer for them
        return ops.dropout(g, None, p, seed, off), (g if ctx.has_res else None), None


class AttentionFn(Function):
    """softmax(QKᵀ/sqrt(hd) + mask)V on packed qkv with optional RoPE and attention-probability dropout —
    model_tiny_gpt.py:94-131."""

    @staticmethod
    def forward(ctx, qkv, seg_start, rope, B, T, H, Hk, hd, window, dropout_p=0.0, bias_masters=None):
        """bias_masters: the (query, key, value) bias Parameters of the packed linear that produced qkv, or None: with
        adjacent flat gradient slots the backward kernels add the column sums of dqkv straight into them."""
        ctx.bias_masters = bias_masters
        if rope is not None:
            ops.rope_qk(qkv, rope[0], rope[1], B, T, H, Hk, hd)  # in place: the QKV GEMM output has no other reader
        seed = off = 0
        if dropout_p > 0.0:
            seed, off = ops.philox_state(qkv.device, 4)  # the mask is indexed by (b,h,i,j); one offset tick per call
        out, lse = ops.attn_fwd(qkv, seg_start, B, T, H, Hk, hd, window=window, dropout_p=dropout_p, seed=seed,
                                offset=off)
        ctx.save_for_backward(qkv, out, lse)
        ctx.aux = (seg_start, rope, B, T, H, Hk, hd, window, dropout_p, seed, off)
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, out, lse = ctx.saved_tensors
        seg_start, rope, B, T, H, Hk, hd, window, dropout_p, seed, off = ctx.aux
        # without RoPE dqkv is final here, so the kernels also take its column sums (the q|k|v bias gradients) and
        # hand them to the QKV linear's backward through the side channel
        csum = None
        direct = False
        if rope is None:
            bm = ctx.bias_masters
            slots = None
            if bm is not None:
                rows, r = [], 0
                for p in bm:
                    rows.append((r, p.shape[0]))
                    r += p.shape[0]
                slots = _packed_slots(bm, rows) if r == (H + 2 * Hk) * hd else None
            if slots is not None:  # the three bias gradients are one contiguous vector of the flat buffer
                csum = torch.as_strided(slots[0], ((H + 2 * Hk) * hd,), (1,))
                direct = True
            else:
                csum = torch.zeros(((H + 2 * Hk) * hd,), dtype=f32, device=qkv.device)
        dqkv = ops.attn_bwd(qkv, seg_start, out, g.contiguous(), lse, B, T, H, Hk, hd, window=window,
                            dropout_p=dropout_p, seed=seed, offset=off, colsum=csum)
        if rope is not None:
            ops.rope_qk(dqkv, rope[0], rope[1], B, T, H, Hk, hd, inverse=True)
        elif direct:
            _credit(bm[0])
        else:
            _attach(dqkv, None, csum)
        return dqkv, None, None, None, None, None, None, None, None, None, None


TC_HEAD_MIN_ROWS = 4096  # below this the fp32 FMA head kernels are used (launch-bound sizes, tests)


_HEAD_W_CACHE = {}  # id(w) -> (w, version, w3, wk): one split per weight and forward pass (6 head calls share the tied head)


def _head_weight_split(w):
    """(w3 [V, 3d] = hi|hi|lo, wk [3Vp, d] = hi;hi;lo stacked along the reduction of the input-gradient GEMM)."""
    hit = _HEAD_W_CACHE.get(id(w))
    if hit is not None and hit[0] is w and hit[1] == w._version:
        return hit[2], hit[3]
    V, d = w.shape
    Vp = (V + 7) // 8 * 8
    with torch.no_grad():
        w3 = ops.split3(w.detach(), partner=True)
        wk = torch.zeros((3 * Vp, d), dtype=bf16, device=w.device)
        wk[0:V], wk[Vp:Vp + V], wk[2 * Vp:2 * Vp + V] = w3[:, 0:d], w3[:, d:2 * d], w3[:, 2 * d:3 * d]
    _HEAD_W_CACHE[id(w)] = (w, w._version, w3, wk)
    return w3, wk


def _split_head_fwd(x, w, bias, x3=None):
    M, d = x.shape
    V = w.shape[0]
    if x3 is None:
        x3 = ops.split3(x)                   # [M, 3d]  hi|lo|hi
    w3, wk = _head_weight_split(w)           # [V, 3d]  hi|hi|lo
    out = torch.empty((M, V), dtype=f32, device=x.device)
    ops.gemm(x3, w3, out, M=M, N=V, K=3 * d, bias=bias)
    return out, x3, w3, wk


def _split_head_bwd(g, x3, w3, wk, wm, bm, dims, need_dx, dx_dtype=f32, dw_into=None):
    """-> (dx or None, dw or None, db or None); dw/db are None when accumulated into main_grad (or `dw_into`)."""
    M, d, V = dims
    Vp = (V + 7) // 8 * 8
    g3 = _grad_form_of(g, (M, 3 * Vp))
    if g3 is None:
        g3 = ops.split3(g.contiguous(), cols_pad=Vp)  # [M, 3Vp]  hi|lo|hi
    dx = None
    if need_dx:
        # dx = g·w: reduction over the (tripled, padded) vocabulary; w as [hi; hi; lo] stacked along K
        dx = torch.empty((M, d), dtype=dx_dtype, device=g.device)
        ops.gemm(g3, wk, dx, M=M, N=d, K=3 * Vp, b_mn=True)
    mw, mb = _main_grad(wm), _main_grad(bm)
    dw = mw if mw is not None else (dw_into if dw_into is not None else torch.zeros((V, d), dtype=f32, device=g.device))
    # dW = gᵀ·x with both operands as [hi|lo]: ONE stacked GEMM gives the four cross products as the quadrants of
    # a [2Vp, 2d] scratch matrix (the tokens are read once, not three times), folded into dW by a small kernel
    scratch = torch.zeros((2 * Vp, 2 * d), dtype=f32, device=g.device)
    tiles = ((2 * Vp + 127) // 128) * ((2 * d + 255) // 256)
    split = ops.pick_split_k(tiles, (M + 63) // 64, _SMS)
    ops.gemm(g3, x3, scratch, M=2 * Vp, N=2 * d, K=M, a_mn=True, b_mn=True, lda=3 * Vp, ldb=3 * d, ldc=2 * d,
             accumulate=True, split_k=split)
    ops.fold_quadrants_add(scratch, dw, V, d, Vp, d)
    db = None
    if bm is not None:
        db = mb if mb is not None else torch.zeros((V,), dtype=f32, device=g.device)
        ops.colsum_bf16(g3, db, N=V, ld=3 * Vp)
        ops.colsum_bf16(g3[:, Vp:], db, N=V, ld=3 * Vp)
    if mw is not None:
        _done(wm)
        if bm is not None:
            _done(bm)
        return dx, None, None
    return dx, (None if dw_into is not None else dw), db


def _plain_head_fwd(hb, w3, V):
    """logits_o = bf16(h_o) · bf16(E)ᵀ with fp32 accumulation, for the OFFSET branches (model_tiny_gpt.py:336).  h_o is
    itself the output of a bf16-operand GEMM (relative error ~2^-9), so the hi/lo split that makes the MAIN head
    fp32-accurate (argmax parity) would only resolve rounding noise here: the shared head is applied to the bf16 h_o
    with the hi part of the split head weight (w3[:, :d], row pitch 3d) — no split pass over h_o, K = d instead of 3d."""
    M, d = hb.shape
    out = torch.empty((M, V), dtype=f32, device=hb.device)
    ops.gemm(hb, w3, out, M=M, N=V, K=d, ldb=w3.stride(0))
    return out


def _plain_head_bwd(g, hb, wk, wm, dims, dw_into=None):
    """-> dh (bf16 [M, d]); the head-weight gradient gᵀ·h_o is accumulated into main_grad / `dw_into` (fp32 [V, d])."""
    M, d, V = dims
    Vp = (V + 7) // 8 * 8
    gb = _grad_form_of(g, (M, Vp))                           # written by the CE backward kernel itself, or:
    if gb is None:
        gb = ops.cast_bf16(g.contiguous(), ld_out=Vp)        # [M, Vp], pad columns zero
    dh = torch.empty((M, d), dtype=bf16, device=g.device)
    ops.gemm(gb, wk, dh, M=M, N=d, K=Vp, b_mn=True)           # wk[0:Vp] = hi rows of the head weight (zero-padded)
    mw = _main_grad(wm)
    dw = mw if mw is not None else dw_into
    tiles = ((V + 127) // 128) * ((d + 255) // 256)
    split = ops.pick_split_k(tiles, (M + 63) // 64, _SMS)
    ops.gemm(gb, hb, dw, M=V, N=d, K=M, a_mn=True, b_mn=True, lda=Vp, ldb=hb.stride(0), ldc=d, accumulate=True,
             split_k=split)
    if mw is not None:
        _done(wm)
    return dh


class HeadsFn(Function):
    """Everything that reads the final hidden state, as ONE autograd node (model_tiny_gpt.py:326-337): LM head,
    termination head and the offset heads (MLP + shared LM head).  The hidden state is split into bf16 hi|lo|hi
    once (its hi part is the offset MLPs' bf16 operand), and in backward every branch ADDS its input gradient to one
    fp32 buffer from inside its own kernel — no cast of the hidden state and no elementwise gradient sums.

    Inputs: x fp32 [M, d], head weight, termination weight/bias (or None), then per offset
    (w1_sh, b1, w2_sh, b2, w1, w2).  Outputs: logits, termination logits (if any), one logits tensor per offset."""

    @staticmethod
    def forward(ctx, x, head_w, term_w, term_b, *off):
        M, d = x.shape
        n_off = len(off) // 6
        need_bwd = any(ctx.needs_input_grad)
        ctx.set_materialize_grads(False)  # an unused branch arrives as None in backward (handled there), not as zeros
        x3 = ops.split3(x)
        logits, _, w3, wk = _split_head_fwd(x, head_w, None, x3=x3)
        outs = [logits]
        if term_w is not None:
            outs.append(ops.skinny_linear_fwd(x, term_w, term_b))
        xb = x3[:, :d]  # bf16(x): row pitch 3d
        saved = [x, x3, w3, wk]
        # All offset MLPs read the same xb.  When their first linears sit back to back in the trainer's flat buffers
        # (TinyGPT.packed_param_groups -> trainer._packed_order) they ARE one [n_off*d, d] operand: one fc1 GEMM here,
        # one fc1 input-gradient GEMM (K = n_off*d, instead of n_off read-modify-write passes over dx) and one fc1
        # weight-gradient GEMM in backward.
        packed = n_off > 1 and _adjacent_tensors([off[6 * o] for o in range(n_off)]) and \
            _adjacent_tensors([off[6 * o + 1].data for o in range(n_off)]) and all(
                off[6 * o].shape == (d, d) for o in range(n_off))
        ctx.packed = bool(packed)
        if packed:
            nd = n_off * d
            w1_all = torch.as_strided(off[0], (nd, d), (d, 1))
            b1_all = torch.as_strided(off[1].data, (nd,), (1,))
            dact_all = torch.empty((M, nd), dtype=bf16, device=x.device) if need_bwd else None
            act_all = torch.empty((M, nd), dtype=bf16, device=x.device)
            ops.gemm(xb, w1_all, act_all, M=M, N=nd, K=d, bias=b1_all, epilogue=EPI_GELU, aux_out=dact_all, ldaux=nd)
        for o in range(n_off):
            w1_sh, b1, w2_sh, b2, w1, w2 = off[6 * o:6 * o + 6]
            if packed:
                act = act_all[:, o * d:(o + 1) * d]
                dact = None if dact_all is None else dact_all[:, o * d:(o + 1) * d]
                hb = torch.empty((M, d), dtype=bf16, device=x.device)
                ops.gemm(act, w2_sh, hb, M=M, N=d, K=d, bias=b2)
            else:
                hb, act, dact = _offset_mlp_fwd(xb, w1_sh, b1, w2_sh, b2, need_bwd, out_dtype=bf16)
            outs.append(_plain_head_fwd(hb, w3, head_w.shape[0]))
            saved += [dact, act, w1_sh, w2_sh, hb]
        if need_bwd:
            ctx.save_for_backward(*saved)
        ctx.masters = (head_w, term_w, term_b) + tuple(off)
        ctx.n_off = n_off
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        saved = ctx.saved_tensors
        x, x3, w3, wk = saved[:4]
        head_w, term_w, term_b = ctx.masters[:3]
        off = ctx.masters[3:]
        M, d = x.shape
        V = head_w.shape[0]
        xb = x3[:, :d]
        g_logits = gs[0]
        g_term = gs[1] if term_w is not None else None
        g_off = gs[(2 if term_w is not None else 1):]
        dhw = None if _main_grad(head_w) is not None else torch.zeros((V, d), dtype=f32, device=x.device)
        if g_logits is not None:
            dx, _, _ = _split_head_bwd(g_logits, x3, w3, wk, head_w, None, (M, d, V), True, f32, dw_into=dhw)
        else:
            dx = torch.zeros((M, d), dtype=f32, device=x.device)
        dtw = dtb = None
        if g_term is not None:
            mw, mb = _main_grad(term_w), _main_grad(term_b)
            dtw = mw if mw is not None else torch.zeros_like(term_w)
            dtb = None
            if term_b is not None:
                dtb = mb if mb is not None else torch.zeros((term_w.shape[0],), dtype=f32, device=x.device)
            ops.skinny_linear_bwd(g_term.contiguous(), x, term_w, dx, True, dtw, dtb)
            if mw is not None:
                _done(term_w)
                if term_b is not None:
                    _done(term_b)
                dtw = dtb = None
        grads = []
        n_off = ctx.n_off
        w1_masters = [off[6 * o + 4] for o in range(n_off)]
        b1_masters = [off[6 * o + 1] for o in range(n_off)]
        rows = [(o * d, d) for o in range(n_off)]
        w1_slots = _packed_slots(w1_masters, rows) if ctx.packed else None
        packed = (ctx.packed and w1_slots is not None and all(g is not None for g in g_off)
                  and all(_main_grad(b) is not None for b in b1_masters))
        if packed:
            # second linears and shared head per offset; the pre-activation gradients of ALL offsets land in one
            # [M, n_off*d] buffer, which then feeds ONE fc1 weight-gradient and ONE fc1 input-gradient GEMM
            nd = n_off * d
            dpre_all = torch.empty((M, nd), dtype=bf16, device=x.device)
            for o in range(n_off):
                dact, act, w1_sh, w2_sh, hb = saved[4 + 5 * o:9 + 5 * o]
                _, b1, _, b2, w1, w2 = off[6 * o:6 * o + 6]
                gb = _plain_head_bwd(g_off[o], hb, wk, head_w, (M, d, V), dw_into=dhw)
                dw2 = _wgrad(gb, act, d, d, master=w2)
                db2 = _colsum(gb, d, master=b2)
                ops.gemm(gb, w2_sh, dpre_all[:, o * d:(o + 1) * d], M=M, N=d, K=d, b_mn=True, epilogue=EPI_MUL_AUX,
                         aux=dact, ldaux=nd, ldc=nd, colsum=_main_grad(b1))
                _done(b1)
                grads += [None, None, None, db2, None, dw2]
            w1_first = saved[4 + 2]  # w1_sh of offset 0: the packed [n_off*d, d] shadow starts there
            _wgrad_into(dpre_all, xb, nd, d, torch.as_strided(w1_slots[0], (nd, d), (d, 1)))
            for m in w1_masters:
                _done(m)
            ops.gemm(dpre_all, torch.as_strided(w1_first, (nd, d), (d, 1)), dx, M=M, N=d, K=nd, b_mn=True, accumulate=True)
            return (dx, dhw, dtw, dtb, *grads)
        for o in range(n_off):
            dact, act, w1_sh, w2_sh, hb = saved[4 + 5 * o:9 + 5 * o]
            _, b1, _, b2, w1, w2 = off[6 * o:6 * o + 6]
            if g_off[o] is None:
                grads += [None] * 6
                continue
            gb = _plain_head_bwd(g_off[o], hb, wk, head_w, (M, d, V), dw_into=dhw)
            _, db1, db2, dw1, dw2 = _offset_mlp_bwd(gb, xb, dact, act, w1_sh, w2_sh, b1, b2, w1, w2, dx_into=dx,
                                                    act_ld=act.stride(0), dact_ld=None if dact is None else dact.stride(0))
            grads += [None, db1, None, db2, dw1, dw2]
        return (dx, dhw, dtw, dtb, *grads)


class SplitHeadFn(Function):
    """The same fp32-accurate head, on tensor cores: x·wᵀ with both operands split into bf16 hi + lo parts and
    the reduction dimension tripled ([hi|lo|hi]·[hi|hi|lo]ᵀ = hi·hi + lo·hi + hi·lo, ~16 mantissa bits), so one
    tcgen05 GEMM replaces the FMA kernel.  Backward uses the same trick for dx (one GEMM) and a stacked
    [hi|lo]ᵀ[hi|lo] GEMM for dW.  LM head :327,336 and termination head :330."""

    @staticmethod
    def forward(ctx, x, w, bias):
        out, x3, w3, wk = _split_head_fwd(x, w, bias)
        ctx.save_for_backward(x3, w3, wk)
        ctx.masters = (w, bias)
        ctx.dims = (x.shape[0], x.shape[1], w.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        x3, w3, wk = ctx.saved_tensors
        wm, bm = ctx.masters
        return _split_head_bwd(g, x3, w3, wk, wm, bm, ctx.dims, ctx.needs_input_grad[0])


class SkinnyLinearFn(Function):
    """fp32 x·wᵀ (+b) for N <= 128 outputs: LM head (:327,336) and termination head (:330)."""

    @staticmethod
    def forward(ctx, x, w, bias):
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.masters = (w, bias)
        return ops.skinny_linear_fwd(x, w, bias)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        wm, bm = ctx.masters
        g = g.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        mw, mb = _main_grad(wm), _main_grad(bm)
        dw = mw if mw is not None else torch.zeros_like(w)
        db = None
        if ctx.has_bias:
            db = mb if mb is not None else torch.zeros((w.shape[0],), dtype=f32, device=w.device)
        ops.skinny_linear_bwd(g, x, w, dx, False, dw, db)
        if mw is not None:
            _done(wm)
            if bm is not None:
                _done(bm)
            return dx, None, None
        return dx, dw, db


class CrossEntropyFn(Function):
    """Mean CE with ignore_index / label smoothing / class weights / offset validity, no host sync
    (model_tiny_gpt.py:343-349; objectives.py:39-57, 94-105).  Returns (loss, kept_weight)."""

    @staticmethod
    def forward(ctx, logits2d, targets, next_boundary, class_w, B, T, shift, smoothing, ignore_index, zero_if_empty,
                grad_form=0):
        # the kernel's last CTA also forms the mean (0 instead of NaN for an empty selection when asked): no torch
        # arithmetic on the sums
        ctx.set_materialize_grads(False)  # no zero tensor for the gradient of `sums`
        sums, row_lse, loss = ops.ce_fwd(logits2d, targets, B, T, shift=shift, next_boundary=next_boundary,
                                         class_w=class_w, smoothing=smoothing, ignore_index=ignore_index,
                                         zero_if_empty=zero_if_empty)
        ctx.save_for_backward(logits2d, row_lse, targets, sums)
        ctx.aux = (next_boundary, class_w, B, T, shift, smoothing, ignore_index, grad_form)
        ctx.mark_non_differentiable(sums)
        return loss, sums

    @staticmethod
    def backward(ctx, g, _gs):
        if g is None:
            return (None,) * 11
        logits2d, row_lse, targets, sums = ctx.saved_tensors
        next_boundary, class_w, B, T, shift, smoothing, ignore_index, grad_form = ctx.aux
        gs = g.reshape(1).to(f32).contiguous()
        # grad_form: the bf16 form of the logit gradient that the head's backward GEMMs read (1 = padded bf16 copy,
        # 2 = hi|lo|hi split), written by the same kernel and handed on ON the gradient tensor (_attach)
        res = ops.ce_bwd(logits2d, row_lse, targets, sums, gs, B, T, shift=shift, next_boundary=next_boundary,
                         class_w=class_w, smoothing=smoothing, ignore_index=ignore_index, bf16_mode=grad_form)
        if grad_form:
            dl, side = res
            _attach(dl, side, None)
        else:
            dl = res
        return dl, None, None, None, None, None, None, None, None, None, None
