"""Tensor-level wrappers over the C ABI: torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import EPI_GELU, EPI_MUL_AUX, EPI_NONE, GemmArgs, check  # noqa: F401

bf16, f32, i32, i64 = torch.bfloat16, torch.float32, torch.int32, torch.int64


def _L():
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.CgptError("cgpt_b200 ops need CUDA tensors: this path has no CPU implementation")
    _lib.ensure_device(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t, dtype, name):
    if t.dtype != dtype or not t.is_contiguous():
        raise _lib.CgptError(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


# ------------------------------------------------------------------ integer scans
def segment_ids(idx: torch.Tensor, sep_id: int) -> torch.Tensor:
    _dev(idx)
    _req(idx, i64, "idx")
    B, T = idx.shape
    out = torch.empty((B, T), dtype=i32, device=idx.device)
    check(_L().cgpt_segment_ids(idx.data_ptr(), out.data_ptr(), B, T, int(sep_id), _stream()))
    return out


def segment_starts(idx: torch.Tensor, sep_id: int) -> torch.Tensor:
    _dev(idx)
    _req(idx, i64, "idx")
    B, T = idx.shape
    out = torch.empty((B, T), dtype=i32, device=idx.device)
    check(_L().cgpt_segment_starts(idx.data_ptr(), out.data_ptr(), B, T, int(sep_id), _stream()))
    return out


def next_in_set(yb: torch.Tensor, ids: Sequence[int]) -> torch.Tensor:
    _dev(yb)
    _req(yb, i64, "yb")
    B, T = yb.shape
    out = torch.empty((B, T), dtype=i32, device=yb.device)
    arr = (C.c_int64 * max(1, len(ids)))(*[int(v) for v in ids])
    check(_L().cgpt_next_in_set(yb.data_ptr(), out.data_ptr(), B, T, arr, len(ids), _stream()))
    return out


def termination_labels(yb: torch.Tensor, next_stop: torch.Tensor, edges: Sequence[int], ignore_index: int = -100):
    _dev(yb)
    B, T = yb.shape
    out = torch.empty((B, T), dtype=i64, device=yb.device)
    arr = (C.c_int64 * max(1, len(edges)))(*[int(v) for v in edges])
    check(_L().cgpt_termination_labels(yb.data_ptr(), next_stop.data_ptr(), out.data_ptr(), B, T, arr, len(edges),
                                       int(ignore_index), _stream()))
    return out


# ------------------------------------------------------------------ embedding
def embed_fwd(idx, tok_w, pos_w):
    _dev(idx)
    B, T = idx.shape
    V, d = tok_w.shape
    x = torch.empty((B, T, d), dtype=f32, device=idx.device)
    check(_L().cgpt_embed_fwd(idx.data_ptr(), tok_w.data_ptr(), _p(pos_w), x.data_ptr(), B, T, d, V, _stream()))
    return x


def embed_bwd(idx, dx, dtok_w, dpos_w):
    B, T = idx.shape
    V, d = dtok_w.shape
    check(_L().cgpt_embed_bwd(idx.data_ptr(), dx.data_ptr(), dtok_w.data_ptr(), _p(dpos_w), B, T, d, V, _stream()))


# ------------------------------------------------------------------ layernorm
def layernorm_fwd(x2d, gamma, beta, want_bf16=True, want_f32=False, eps=1e-5):
    _dev(x2d)
    M, d = x2d.shape
    yb = torch.empty((M, d), dtype=bf16, device=x2d.device) if want_bf16 else None
    yf = torch.empty((M, d), dtype=f32, device=x2d.device) if want_f32 else None
    mean = torch.empty((M,), dtype=f32, device=x2d.device)
    rstd = torch.empty((M,), dtype=f32, device=x2d.device)
    check(_L().cgpt_layernorm_fwd(x2d.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(yb), _p(yf), mean.data_ptr(),
                                  rstd.data_ptr(), M, d, float(eps), _stream()))
    return yb, yf, mean, rstd


def layernorm_bwd(dy, x2d, gamma, mean, rstd, dres, dgamma, dbeta, want_bf16=False, dx_colsum=None):
    M, d = x2d.shape
    dx = torch.empty((M, d), dtype=f32, device=x2d.device)
    dxb = torch.empty((M, d), dtype=bf16, device=x2d.device) if want_bf16 else None
    check(_L().cgpt_layernorm_bwd(dy.data_ptr(), 1 if dy.dtype == f32 else 0, x2d.data_ptr(), gamma.data_ptr(),
                                  mean.data_ptr(), rstd.data_ptr(), _p(dres), dx.data_ptr(), _p(dxb),
                                  dgamma.data_ptr(), dbeta.data_ptr(), _p(dx_colsum), M, d, _stream()))
    return dx, dxb


# ------------------------------------------------------------------ GEMM
def gemm(a, b, out, *, M, N, K, a_mn=False, b_mn=False, lda=None, ldb=None, ldc=None, bias=None, epilogue=EPI_NONE,
         aux=None, aux_out=None, ldaux=0, residual=None, accumulate=False, split_k=1, colsum=None):
    """out[M,N] (+)= A·Bᵀ (+bias, epilogue).  a/b bf16; out bf16 or fp32 (by dtype)."""
    g = GemmArgs()
    g.a, g.b = a.data_ptr(), b.data_ptr()
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.lda = int(lda if lda is not None else a.stride(0))
    g.ldb = int(ldb if ldb is not None else b.stride(0))
    g.M, g.N, g.K = int(M), int(N), int(K)
    g.split_k = int(split_k)
    g.bias = _p(bias)
    g.epilogue = int(epilogue)
    g.aux, g.aux_out, g.ldaux = _p(aux), _p(aux_out), int(ldaux)
    g.residual = _p(residual)
    g.out = out.data_ptr()
    g.out_f32 = 1 if out.dtype == f32 else 0
    g.accumulate = int(accumulate)
    g.ldc = int(ldc if ldc is not None else out.stride(0))
    g.colsum = _p(colsum)
    check(_L().cgpt_gemm_bf16(C.byref(g), _stream()))
    return out


def pick_split_k(tiles: int, kblocks: int, sms: int = 148, max_split: int = 24) -> int:
    """Split the reduction of a weight-gradient GEMM so that (tiles x split) fills whole waves of the
    persistent grid: maximise wave efficiency, prefer fewer splits (less atomic traffic) on ties."""
    if tiles >= 2 * sms or kblocks <= 8:
        return 1
    best, best_eff = 1, 0.0
    for s in range(1, max(1, min(max_split, kblocks // 8)) + 1):
        work = tiles * s
        eff = work / (-(-work // sms) * sms)
        if eff > best_eff + 0.02:
            best, best_eff = s, eff
    return best


# ------------------------------------------------------------------ elementwise
def cast_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None, ld_out: Optional[int] = None):
    """fp32 [rows, cols] -> bf16 [rows, ld_out] (pad columns zeroed)."""
    _dev(src)
    s2 = src if src.dim() == 2 else src.reshape(1, -1)
    rows, cols = s2.shape
    ld_out = cols if ld_out is None else ld_out
    if out is None:
        out = torch.empty((rows, ld_out), dtype=bf16, device=src.device)
    check(_L().cgpt_cast_f32_bf16(s2.data_ptr(), s2.stride(0), out.data_ptr(), ld_out, rows, cols, _stream()))
    return out


def split3(src: torch.Tensor, partner: bool = False, cols_pad: Optional[int] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [rows, 3*cols_pad]: [hi|lo|hi] (or [hi|hi|lo] for the partner operand)."""
    _dev(src)
    rows, cols = src.shape
    cols_pad = (cols + 7) // 8 * 8 if cols_pad is None else cols_pad
    out = torch.empty((rows, 3 * cols_pad), dtype=bf16, device=src.device)
    check(_L().cgpt_split3_f32_bf16(src.data_ptr(), src.stride(0), out.data_ptr(), rows, cols, cols_pad, int(partner),
                                    _stream()))
    return out


def fold_quadrants_add(s: torch.Tensor, dst: torch.Tensor, rows: int, cols: int, row_off: int, col_off: int):
    """dst[r,c] += s[r,c] + s[r,c+col_off] + s[r+row_off,c] + s[r+row_off,c+col_off]  (fp32)."""
    _dev(s)
    _req(s, f32, "s")
    _req(dst, f32, "dst")
    check(_L().cgpt_fold_quadrants_add(s.data_ptr(), s.stride(0), dst.data_ptr(), dst.stride(0), rows, cols, row_off,
                                       col_off, _stream()))
    return dst


def colsum_bf16(x2d, out, N=None, ld=None):
    M = x2d.shape[0]
    check(_L().cgpt_colsum_bf16(x2d.data_ptr(), int(ld if ld is not None else x2d.stride(0)), out.data_ptr(), M,
                                int(N if N is not None else x2d.shape[1]), _stream()))


def rope_qk(qkv, cos_t, sin_t, B, T, H, Hk, hd, inverse=False):
    check(_L().cgpt_rope_qk(qkv.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), B, T, H, Hk, hd, int(inverse), _stream()))


def swiglu_fwd(gu, hp):
    M = gu.shape[0]
    act = torch.empty((M, hp), dtype=bf16, device=gu.device)
    check(_L().cgpt_swiglu_fwd(gu.data_ptr(), gu.stride(0), act.data_ptr(), hp, M, hp, _stream()))
    return act


def swiglu_bwd(gu, dact, hp):
    M = gu.shape[0]
    dgu = torch.empty_like(gu)
    check(_L().cgpt_swiglu_bwd(gu.data_ptr(), gu.stride(0), dact.data_ptr(), dact.stride(0), dgu.data_ptr(), M, hp,
                               _stream()))
    return dgu


# ------------------------------------------------------------------ attention
class DeviceGenerator:
    """Dropout generator state that lives on the device: {seed, base offset} as two int64 in HBM
    (cgpt_set_philox_state).  Kernels add the base to their by-value offset themselves, so a CUDA graph that
    captured a training step with dropout draws new masks on every replay; `advance()` (one launch, captured with
    the step) moves the base past everything the step consumed.  Seeded from torch's CUDA generator, so
    torch.manual_seed still governs the masks (reference tests/test_attention_dropout.py:62-78)."""

    def __init__(self, device: torch.device):
        gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
        seed = gen.initial_seed() & 0x7FFFFFFFFFFFFFFF
        self.state = torch.tensor([seed, gen.get_offset()], dtype=torch.int64, device=device)
        self.consumed = 0  # by-value offsets handed out since the last advance()

    def draw(self, n_random: int):
        off = self.consumed
        self.consumed += (int(n_random) + 3) // 4 * 4
        return 0, off

    def advance(self):
        """Move the base offset past the draws made since the previous advance (device-side add)."""
        if self.consumed:
            check(_L().cgpt_philox_advance(self.state.data_ptr(), self.consumed, _stream()))
        inc, self.consumed = self.consumed, 0
        return inc


_ACTIVE_GEN = [None]


class device_generator:
    """Context: dropout kernels launched inside read seed / base offset from `gen.state` (forward AND the backward
    kernels that autograd launches from its own thread: the switch is process-wide)."""

    def __init__(self, gen: Optional["DeviceGenerator"]):
        self.gen = gen

    def __enter__(self):
        self.prev = _ACTIVE_GEN[0]
        _ACTIVE_GEN[0] = self.gen
        check(_L().cgpt_set_philox_state(None if self.gen is None else self.gen.state.data_ptr()))
        return self.gen

    def __exit__(self, *exc):
        _ACTIVE_GEN[0] = self.prev
        check(_L().cgpt_set_philox_state(None if self.prev is None else self.prev.state.data_ptr()))
        return False


def philox_state(device: torch.device, n_random: int):
    """(seed, offset) for a dropout kernel.  Default: torch's CUDA generator, advanced by n_random on the host —
    torch.manual_seed governs dropout (reference tests/test_attention_dropout.py:62-78) without any kernel launch or
    host sync.  Inside a `device_generator` scope: (0, offset within the step); the kernels add the device-resident
    seed and base offset."""
    if _ACTIVE_GEN[0] is not None:
        return _ACTIVE_GEN[0].draw(n_random)
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, off = gen.initial_seed(), gen.get_offset()
    gen.set_offset(off + (int(n_random) + 3) // 4 * 4)
    return seed & 0xFFFFFFFFFFFFFFFF, off


def dropout(x: torch.Tensor, residual: Optional[torch.Tensor], p: float, seed: int, offset: int, out_bf16=False):
    """out = (residual) + x * mask / (1-p); x fp32, numel % 4 == 0."""
    _dev(x)
    out = torch.empty(x.shape, dtype=bf16 if out_bf16 else f32, device=x.device)
    check(_L().cgpt_dropout(x.data_ptr(), _p(residual), out.data_ptr(), int(out_bf16), x.numel(), float(p), seed, offset,
                            _stream()))
    return out


def _pad_heads(x, n_heads, hd, hdp):
    """[M, n_heads*hd] -> [M, n_heads*hdp], zeros in the extra columns of every head."""
    M = x.shape[0]
    out = torch.zeros((M, n_heads, hdp), dtype=x.dtype, device=x.device)
    out[:, :, :hd] = x.reshape(M, n_heads, hd)
    return out.view(M, n_heads * hdp)


def _unpad_heads(x, n_heads, hd, hdp):
    return x.view(x.shape[0], n_heads, hdp)[:, :, :hd].reshape(x.shape[0], n_heads * hd).contiguous()


def _padded_hd(hd):
    """Head sizes the tensor-core kernels do not tile (not a multiple of 16: the reference's own unit tests use
    n_embd 8 / 16 with 2 heads, tests/test_attention_dropout.py:11-12, test_embedding_extraction_contract.py:18-19) run
    on zero-padded heads: zero q/k columns add nothing to QKᵀ, zero v columns give zero output columns, and the
    softmax scale stays 1/sqrt(hd) of the TRUE width.  A compatibility path for toy models, not a hot path."""
    return hd if hd % 16 == 0 else (hd + 15) // 16 * 16


def attn_fwd(qkv, seg_start, B, T, H, Hk, hd, window=0, scale=None, dropout_p=0.0, seed=0, offset=0):
    _dev(qkv)
    scale = 1.0 / math.sqrt(hd) if scale is None else scale
    hdp = _padded_hd(hd)
    if hdp != hd:
        out, lse = attn_fwd(_pad_heads(qkv, H + 2 * Hk, hd, hdp), seg_start, B, T, H, Hk, hdp, window, scale, dropout_p,
                            seed, offset)
        return _unpad_heads(out, H, hd, hdp), lse
    out = torch.empty((B * T, H * hd), dtype=bf16, device=qkv.device)
    lse = torch.empty((B, H, T), dtype=f32, device=qkv.device)
    check(_L().cgpt_attn_fwd(qkv.data_ptr(), _p(seg_start), out.data_ptr(), lse.data_ptr(), B, T, H, Hk, hd,
                             int(window or 0), float(scale), float(dropout_p), seed, offset, _stream()))
    return out, lse


def attn_bwd(qkv, seg_start, out, dout, lse, B, T, H, Hk, hd, window=0, scale=None, dropout_p=0.0, seed=0, offset=0,
             colsum=None):
    """dqkv; `colsum` (fp32 [(H+2Hk)*hd], accumulated) also receives the column sums of dqkv (q|k|v bias grads)."""
    scale = 1.0 / math.sqrt(hd) if scale is None else scale
    hdp = _padded_hd(hd)
    if hdp != hd:
        dq = attn_bwd(_pad_heads(qkv, H + 2 * Hk, hd, hdp), seg_start, _pad_heads(out, H, hd, hdp),
                      _pad_heads(dout, H, hd, hdp), lse, B, T, H, Hk, hdp, window, scale, dropout_p, seed, offset)
        dq = _unpad_heads(dq, H + 2 * Hk, hd, hdp)
        if colsum is not None:
            colsum.add_(dq.float().sum(0))
        return dq
    dqkv = torch.empty_like(qkv)
    nbytes = _L().cgpt_attn_bwd_workspace(B, T, H, Hk, hd)
    ws = torch.empty((nbytes // 4,), dtype=f32, device=qkv.device)
    check(_L().cgpt_attn_bwd_colsum(qkv.data_ptr(), _p(seg_start), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                    dqkv.data_ptr(), ws.data_ptr(), _p(colsum), B, T, H, Hk, hd, int(window or 0),
                                    float(scale), float(dropout_p), seed, offset, _stream()))
    return dqkv


def attn_probs(qkv, seg_start, B, T, H, Hk, hd, window=0, scale=None, dropout_p=0.0, seed=0, offset=0):
    scale = 1.0 / math.sqrt(hd) if scale is None else scale
    hdp = _padded_hd(hd)
    if hdp != hd:
        return attn_probs(_pad_heads(qkv, H + 2 * Hk, hd, hdp), seg_start, B, T, H, Hk, hdp, window, scale, dropout_p,
                          seed, offset)
    att = torch.empty((B, H, T, T), dtype=f32, device=qkv.device)
    check(_L().cgpt_attn_probs(qkv.data_ptr(), _p(seg_start), att.data_ptr(), B, T, H, Hk, hd, int(window or 0),
                               float(scale), float(dropout_p), seed, offset, _stream()))
    return att


def attn_decode(qkv_new, k_cache, v_cache, lo, t, H, Hk, hd, window=0, scale=None, t_dev=None):
    """One new token per sequence against the (K, V) cache [B, Tmax, Hk*hd]; appends position t. -> bf16 [B, H*hd].
    t_dev (int32 device scalar) overrides t: the position is then read on the device (graph replay)."""
    _dev(qkv_new)
    B = qkv_new.shape[0]
    scale = 1.0 / math.sqrt(hd) if scale is None else scale
    out = torch.empty((B, H * hd), dtype=bf16, device=qkv_new.device)
    check(_L().cgpt_attn_decode(qkv_new.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), _p(lo), out.data_ptr(), B,
                                int(t), _p(t_dev), k_cache.shape[1], H, Hk, hd, int(window or 0), float(scale),
                                _stream()))
    return out


def pack_lm_batch(tokens, offsets, lengths, indices, T_out):
    """(xb, yb) int64 [B, T_out] from the device-resident packed dataset for the sequence `indices` (int64, device)."""
    _dev(tokens)
    B = indices.numel()
    xb = torch.empty((B, T_out), dtype=torch.int64, device=tokens.device)
    yb = torch.empty((B, T_out), dtype=torch.int64, device=tokens.device)
    check(_L().cgpt_pack_lm_batch(tokens.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), indices.data_ptr(), B,
                                  int(T_out), xb.data_ptr(), yb.data_ptr(), _stream()))
    return xb, yb


def shape_proj_fwd(x2d, s2d, w, b):
    """x + shape_embeddings·wᵀ + b  (fp32; x [M, d], shape_embeddings [M, 3], w [d, 3])."""
    _dev(x2d)
    M, d = x2d.shape
    out = torch.empty_like(x2d)
    check(_L().cgpt_shape_proj_fwd(x2d.data_ptr(), s2d.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), M, d,
                                   _stream()))
    return out


def shape_proj_bwd(dx, s2d, w, dw, db, want_ds):
    M, d = dx.shape
    ds = torch.empty((M, 3), dtype=f32, device=dx.device) if want_ds else None
    check(_L().cgpt_shape_proj_bwd(dx.data_ptr(), s2d.data_ptr(), w.data_ptr(), dw.data_ptr(), db.data_ptr(), _p(ds),
                                   M, d, _stream()))
    return ds


# ------------------------------------------------------------------ heads / loss
def skinny_linear_fwd(x2d, w, bias=None):
    _dev(x2d)
    M, d = x2d.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=f32, device=x2d.device)
    check(_L().cgpt_skinny_linear_fwd(x2d.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), M, N, d, _stream()))
    return out


def skinny_linear_bwd(dout, x2d, w, dx, dx_accumulate, dw, dbias):
    M, d = x2d.shape
    N = w.shape[0]
    check(_L().cgpt_skinny_linear_bwd(dout.data_ptr(), x2d.data_ptr(), w.data_ptr(), _p(dx), int(dx_accumulate),
                                      _p(dw), _p(dbias), M, N, d, _stream()))


def ce_fwd(logits2d, targets, B, T, *, shift=0, next_boundary=None, class_w=None, smoothing=0.0, ignore_index=0,
           zero_if_empty=False):
    """Returns (sums[2] = (loss_sum, weight_sum), row_lse[M], mean loss (0-dim)); the mean is 0 instead of NaN for an
    empty selection when `zero_if_empty`."""
    _dev(logits2d)
    M, V = logits2d.shape
    sums = torch.empty((2,), dtype=f32, device=logits2d.device)
    mean = torch.empty((), dtype=f32, device=logits2d.device)
    row_lse = torch.empty((M,), dtype=f32, device=logits2d.device)
    row_ws = torch.empty((2 * M,), dtype=f32, device=logits2d.device)
    check(_L().cgpt_ce_fwd(logits2d.data_ptr(), targets.data_ptr(), _p(next_boundary), _p(class_w), sums.data_ptr(),
                           mean.data_ptr(), row_lse.data_ptr(), row_ws.data_ptr(), B, T, V, int(shift), float(smoothing),
                           int(ignore_index), int(bool(zero_if_empty)), _stream()))
    return sums, row_lse, mean


def ce_bwd(logits2d, row_lse, targets, sums, gscale, B, T, *, coef=1.0, shift=0, next_boundary=None, class_w=None,
           smoothing=0.0, ignore_index=0, bf16_mode=0):
    """-> dlogits (fp32), or (dlogits, by-product) when bf16_mode != 0: 1 = bf16 [M, Vp] copy (pad columns zero),
    2 = bf16 [M, 3*Vp] hi|lo|hi split, Vp = V rounded up to 8 — what the head GEMMs of the backward pass read."""
    M, V = logits2d.shape
    dlogits = torch.empty((M, V), dtype=f32, device=logits2d.device)
    Vp = (V + 7) // 8 * 8
    side = None
    if bf16_mode:
        side = torch.empty((M, (3 if bf16_mode == 2 else 1) * Vp), dtype=bf16, device=logits2d.device)
    check(_L().cgpt_ce_bwd(logits2d.data_ptr(), row_lse.data_ptr(), targets.data_ptr(), _p(next_boundary), _p(class_w),
                           sums.data_ptr(), _p(gscale), float(coef), dlogits.data_ptr(), _p(side), int(bf16_mode), Vp,
                           B, T, V, int(shift), float(smoothing), int(ignore_index), _stream()))
    return dlogits if not bf16_mode else (dlogits, side)


# ------------------------------------------------------------------ optimiser
def adamw(p, g, m, v, shadow, lr, beta1, beta2, eps, wd, step, grad_scale=1.0, dev_hyper=None):
    """Fused AdamW on flat buffers; `g` fp32, or bf16 (the all-reduced gradient buckets, consumed in place)."""
    fn = _L().cgpt_adamw_bf16grad if g.dtype == bf16 else _L().cgpt_adamw
    check(fn(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(shadow), p.numel(), float(lr),
                          float(beta1), float(beta2), float(eps), float(wd), int(step), float(grad_scale),
                          _p(dev_hyper), _stream()))


def launch_count() -> int:
    return int(_L().cgpt_launch_count())
