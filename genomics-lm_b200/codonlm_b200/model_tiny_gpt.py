"""Drop-in for the reference's ``src/codonlm/model_tiny_gpt.py`` running on libcgpt_b200 (sm_100a).

Same classes, constructor arguments, attribute / submodule names, ``state_dict`` keys and shapes,
``forward(idx, targets, return_aux, shape_embeddings, attention_window)`` contract and error
behaviour as the reference (model_tiny_gpt.py:9-389); submodules are created in the reference's order
so ``torch.manual_seed(s); TinyGPT(...)`` yields bit-identical initial weights.

What differs is where the arithmetic happens: every op runs in a hand-written CUDA kernel behind the
C ABI of include/cgpt.h (fp32 residual stream and LayerNorm statistics, bf16 tensor-core GEMMs and
attention with fp32 accumulation, fp32 LM head and cross-entropy).  There is no CPU path: calling the
model with parameters that are not on a B200 raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from . import functional as Fn
from .ops import bf16, f32


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
@dataclass
class MaskSpec:
    """The reference's boolean (B,1,T,T) mask (model_tiny_gpt.py:273-295) in the form the kernels use:
    allowed[i,j] = seg_start[b,i] <= j <= i and i-j < window."""
    seg_start: Optional[torch.Tensor]  # int32 [B,T] or None (pure causal)
    window: int = 0                    # 0 = unlimited


def mask_spec_from_bool(mask: torch.Tensor, B: int, T: int) -> MaskSpec:
    """The reference hands its attention a boolean mask tensor (B,1,T,T) / (1,T,T) / (T,T), True = may attend
    (model_tiny_gpt.py:106-113, built by build_attention_mask :273-295).  Every mask that function can build is a per-row
    interval [lo_i, i]; it is converted back to that form (first allowed column per row) and checked — one host
    sync, on this compatibility path only.  Masks that are not causal intervals are refused."""
    m = mask
    if m.dim() == 2:
        m = m[None, None]
    elif m.dim() == 3:
        m = m[:, None]
    if m.dim() != 4 or m.shape[1] != 1 or m.shape[-1] != T or m.shape[-2] != T or m.shape[0] not in (1, B):
        raise NotImplementedError(f"attention mask of shape {tuple(mask.shape)}: expected (B,1,T,T), (1,T,T) or (T,T)")
    m = m.to(torch.bool).expand(B, 1, T, T)[:, 0]
    as_int = m.to(torch.int8)
    lo = as_int.argmax(dim=-1)
    hi = (T - 1) - as_int.flip(-1).argmax(dim=-1)
    rows = torch.arange(T, device=m.device)[None, :]
    ok = (hi == rows) & (as_int.sum(-1) == rows - lo + 1)
    if not bool(ok.all()):
        raise NotImplementedError("codonlm_b200 attention supports causal interval masks (causal / same-segment / "
                                  "window, as build_attention_mask produces them); this mask is not one")
    return MaskSpec(lo.to(torch.int32).contiguous(), 0)


class _CastBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x.reshape(-1, x.shape[-1]).contiguous()).view(x.shape)

    @staticmethod
    def backward(ctx, g):
        return g.float()


def _has_hooks(m: nn.Module) -> bool:
    return bool(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.CgptError(
            f"{what} is on '{t.device}': codonlm_b200 has no CPU implementation — move the model and its "
            "inputs to a B200 (model.to('cuda'))")


# ----------------------------------------------------------------------------------------------
# host staging (SURVEY hard part 6): reference callers that keep the model on the CPU
# ----------------------------------------------------------------------------------------------
_HOOK_TABLES = ("_forward_hooks", "_forward_pre_hooks", "_forward_hooks_with_kwargs",
                "_forward_pre_hooks_with_kwargs", "_forward_hooks_always_called")


def _staging_device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.CgptError(
            "codonlm_b200 has no CPU implementation and no CUDA device is visible: the model's parameters are on the "
            "host, and staging them needs a B200")
    return torch.device("cuda", torch.cuda.current_device())


def _device_twin(mod: nn.Module) -> nn.Module:
    """CUDA replica of a CPU-resident module.  The reference's inference callers pick their device as
    `mps if available else cpu` (scripts/query_model.py:29-34, src/codonlm/score_mutations.py:29-31) and therefore keep
    the model and its inputs on the host; the replica lets them run unchanged: it is built on first use, refreshed
    when a host parameter / buffer changes (load_state_dict, in-place edits: tensor._version), shares the host
    module's forward-hook tables (hooks registered on the host modules fire, with device tensors), and mirrors
    train / eval flags and `use_sdpa`.  Results are copied back to the caller's device by the call sites."""
    import copy
    dev = _staging_device()
    tensors = list(mod.parameters()) + list(mod.buffers())
    key = tuple(t._version for t in tensors) + tuple(id(t) for t in tensors)
    twin = mod.__dict__.get("_cgpt_twin")
    if twin is None or mod.__dict__.get("_cgpt_twin_dev") != dev:
        saved = {k: mod.__dict__.pop(k) for k in ("_cgpt_twin", "_cgpt_twin_key", "_cgpt_twin_dev") if k in mod.__dict__}
        try:
            twin = copy.deepcopy(mod).to(dev)
        finally:
            mod.__dict__.update(saved)
        for a, b in zip(mod.modules(), twin.modules()):
            for name in _HOOK_TABLES:
                if name in a.__dict__:
                    b.__dict__[name] = a.__dict__[name]  # the SAME tables: later (de)registrations are seen
        mod.__dict__["_cgpt_twin"], mod.__dict__["_cgpt_twin_dev"] = twin, dev
        mod.__dict__["_cgpt_twin_key"] = key
    elif mod.__dict__.get("_cgpt_twin_key") != key:
        with torch.no_grad():
            for src, dst in zip(tensors, list(twin.parameters()) + list(twin.buffers())):
                dst.copy_(src)
        mod.__dict__["_cgpt_twin_key"] = key
    for a, b in zip(mod.modules(), twin.modules()):
        b.training = a.training
        if hasattr(a, "use_sdpa"):
            b.use_sdpa = a.use_sdpa
    return twin


def _to_device(obj, dev):
    if isinstance(obj, torch.Tensor):
        return obj.to(dev)
    if isinstance(obj, dict):
        return {k: _to_device(v, dev) for k, v in obj.items()}
    if isinstance(obj, (tuple, list)):
        return type(obj)(_to_device(v, dev) for v in obj)
    return obj


def _host_call(mod: nn.Module, method: str, args, kwargs, home: torch.device):
    """Run `mod.<method>` on the CUDA replica of a host-resident module and bring the results home.  Inference only:
    a host-resident model cannot be TRAINED through the replica (its gradients would live on another module)."""
    _staging_device()  # fails loudly without a CUDA device: there is no CPU path
    if torch.is_grad_enabled() and mod.training and any(p.requires_grad for p in mod.parameters()):
        raise _lib.CgptError(
            "training a host-resident model: move it to the GPU first (model.to('cuda')), as the reference trainer does "
            "(loop.py:486-579); host staging serves inference callers only")
    twin = _device_twin(mod)
    dev = next(twin.parameters()).device
    with torch.no_grad():
        out = getattr(twin, method)(*_to_device(args, dev), **_to_device(kwargs, dev))
    for a, b in zip(mod.modules(), twin.modules()):
        if isinstance(a, CausalSelfAttention):
            la = b.last_attn
            a.last_attn = None if la is None else la.to(home)
    return _to_device(out, home)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


# ----------------------------------------------------------------------------------------------
# leaf modules (same parameters / keys as nn.LayerNorm, nn.Embedding, nn.Linear)
# ----------------------------------------------------------------------------------------------
class LayerNorm(nn.LayerNorm):
    def forward(self, x, colsum_target=None):  # public path: fp32 in -> fp32 out (hooks on ln_f see what the reference shows)
        _require_cuda(self.weight, "LayerNorm.weight")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).float().contiguous()
        _, _, yf = Fn.ResidualLayerNormFn.apply(x2, self.weight, self.bias, True, colsum_target)
        return yf.view(shp)

    def fused(self, x2d, colsum_target=None):
        """(residual alias, bf16 normalised) for the pre-norm residual pattern.  colsum_target: the bias of the
        residual linear that produced x2d (its gradient = column sums of this LayerNorm's input gradient)."""
        return Fn.ResidualLayerNormFn.apply(x2d, self.weight, self.bias, False, colsum_target)


class Embedding(nn.Embedding):
    def forward(self, idx):
        _require_cuda(self.weight, "Embedding.weight")
        flat = idx.to(self.weight.device).long().reshape(1, -1).contiguous()
        return Fn.EmbedFn.apply(flat, self.weight, None).view(*idx.shape, -1)


class Linear(nn.Linear):
    """nn.Linear whose forward is the tcgen05 GEMM (bf16 operands, fp32 accumulate, fp32 output)."""

    def forward(self, x):
        _require_cuda(self.weight, "Linear.weight")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        xb = x2 if x2.dtype == bf16 else _CastBf16.apply(x2.float())
        n, k = self.weight.shape
        if k % 8 != 0:
            raise _lib.CgptError(f"Linear in_features={k} must be a multiple of 8 for the TMA path")
        w_sh = ops.cast_bf16(self.weight.detach())
        masters = (self.weight,) if self.bias is None else (self.weight, self.bias)
        out = Fn.PackedLinearFn.apply(xb.contiguous(), w_sh, None if self.bias is None else self.bias.detach(), None,
                                      ((0, n),), True, *masters)
        return out.view(*shp[:-1], n)


class SkinnyLinear(nn.Linear):
    """fp32 head with <= 128 outputs (LM head, termination head)."""

    def forward(self, x):
        _require_cuda(self.weight, "head.weight")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).float().contiguous()
        if self.out_features > 128:
            raise _lib.CgptError(f"head with {self.out_features} outputs > 128 is not supported by the fp32 head kernel")
        use_tc = (x2.shape[0] >= Fn.TC_HEAD_MIN_ROWS and shp[-1] % 8 == 0 and self.out_features % 4 == 0)
        fn = Fn.SplitHeadFn if use_tc else Fn.SkinnyLinearFn
        out = fn.apply(x2, self.weight, self.bias)
        return out.view(*shp[:-1], self.out_features)


class RotaryEmbedding(nn.Module):
    """cos/sin tables exactly as the reference builds them (model_tiny_gpt.py:9-33)."""

    def __init__(self, dim, max_position_embeddings=512, base=10000, device=None):
        super().__init__()
        self.dim = dim
        self.max_position_embeddings = max_position_embeddings
        self.base = base
        inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.float32) / self.dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self._set_cos_sin_cache(seq_len=max_position_embeddings, device=device, dtype=torch.get_default_dtype())

    def _set_cos_sin_cache(self, seq_len, device, dtype):
        self.max_seq_len_cached = seq_len
        t = torch.arange(self.max_seq_len_cached, device=device, dtype=self.inv_freq.dtype)
        freqs = torch.outer(t, self.inv_freq.to(t.device))
        emb = torch.cat((freqs, freqs), dim=-1)
        self.register_buffer("cos_cached", emb.cos().to(dtype), persistent=False)
        self.register_buffer("sin_cached", emb.sin().to(dtype), persistent=False)
        self._half = None

    def forward(self, x, seq_len=None):
        if seq_len > self.max_seq_len_cached:
            self._set_cos_sin_cache(seq_len=seq_len, device=x.device, dtype=torch.float32)
        return self.cos_cached[:seq_len].to(x.device), self.sin_cached[:seq_len].to(x.device)

    def half_tables(self, T, device):
        """fp32 [T, dim/2] cos / sin (the two halves of the reference tables are identical)."""
        if T > self.max_seq_len_cached:
            self._set_cos_sin_cache(seq_len=T, device=device, dtype=torch.float32)
        if self._half is None or self._half[0].device != device or self._half[0].shape[0] < T:
            h = self.dim // 2
            self._half = (self.cos_cached[:, :h].to(device=device, dtype=f32).contiguous(),
                          self.sin_cached[:, :h].to(device=device, dtype=f32).contiguous())
        return self._half


def apply_rotary_pos_emb(q, k, cos, sin):
    cos = cos.unsqueeze(0).unsqueeze(1)
    sin = sin.unsqueeze(0).unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------
_SHADOW_GEN = [0]


def bump_shadow_generation():
    """Call after writing master weights outside autograd's view (e.g. the fused AdamW kernel)."""
    _SHADOW_GEN[0] += 1
    Fn.reset_head_cache()


def flat_shadow(p):
    """bf16 view of parameter `p` inside the trainer's flat shadow buffer (trainer.FlatGroup), or None when the
    model is not driven by TrainStep.  The fused AdamW kernel keeps it current; a master that was written through
    autograd-visible ops since (load_state_dict, manual edits: its _version moved) is re-cast here."""
    sh = getattr(p, "_cgpt_shadow", None)
    if sh is None or not p.is_cuda:
        return None
    if p._version != p._cgpt_shadow_version:
        with torch.no_grad():
            ops.cast_bf16(p.detach().reshape(1, -1), out=sh.view(1, -1))
        p._cgpt_shadow_version = p._version
    return sh


def _adjacent(tensors):
    """True when the tensors sit back to back in memory (same dtype), i.e. form one packed row-major matrix."""
    for a, b in zip(tensors[:-1], tensors[1:]):
        if a.dtype != b.dtype or a.data_ptr() + a.numel() * a.element_size() != b.data_ptr():
            return False
    return True


class _ShadowMixin:
    """bf16 (and packed) copies of the fp32 master parameters, rebuilt when a master changes."""

    def _shadow_key(self, params):
        return (_SHADOW_GEN[0],) + tuple((p.data_ptr(), p._version) for p in params)

    def _get_shadow(self, name, params, builder):
        cache = self.__dict__.setdefault("_shadow_cache", {})
        key = self._shadow_key(params)
        hit = cache.get(name)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, builder())
            cache[name] = hit
        return hit[1]


class SwiGLU(nn.Module, _ShadowMixin):
    def __init__(self, n_embd, dropout):
        super().__init__()
        hidden_dim = int(8 * n_embd // 3)
        self.w_gate = nn.Linear(n_embd, hidden_dim, bias=False)
        self.w_up = nn.Linear(n_embd, hidden_dim, bias=False)
        self.w_down = nn.Linear(hidden_dim, n_embd, bias=False)
        self.dropout = nn.Dropout(dropout)

    def _shadows(self):
        h, d = self.w_gate.weight.shape
        hp = (h + 7) // 8 * 8

        def build():
            wgu = torch.zeros((2 * hp, d), dtype=bf16, device=self.w_gate.weight.device)
            ops.cast_bf16(self.w_gate.weight, out=wgu[:h])
            ops.cast_bf16(self.w_up.weight, out=wgu[hp:hp + h])
            wd = ops.cast_bf16(self.w_down.weight, ld_out=hp)
            return wgu, wd
        return self._get_shadow("swiglu", (self.w_gate.weight, self.w_up.weight, self.w_down.weight), build)

    def forward(self, x, residual=None):
        _require_cuda(self.w_gate.weight, "SwiGLU weights")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        xb = x2 if x2.dtype == bf16 else _CastBf16.apply(x2.float())
        wgu, wd = self._shadows()
        r2 = None if residual is None else residual.reshape(-1, shp[-1])
        drop = self.training and self.dropout.p > 0.0
        out = Fn.MlpSwiGLUFn.apply(xb.contiguous(), None if drop else r2, wgu, wd, self.w_gate.weight, self.w_up.weight,
                                   self.w_down.weight)
        if drop:  # dropout sits between the MLP and the residual add (:57,152)
            out = Fn.DropoutFn.apply(out, r2, float(self.dropout.p))
        return out.view(shp[:-1] + (out.shape[-1],))


class GeluMLP(nn.Sequential, _ShadowMixin):
    """nn.Sequential(Linear(d,4d), GELU(), Linear(4d,d), Dropout) with the reference's child indices
    (state_dict keys mlp.0.*, mlp.2.*), evaluated as two fused GEMMs."""

    def __init__(self, n_embd, dropout):
        super().__init__(nn.Linear(n_embd, 4 * n_embd), nn.GELU(), nn.Linear(4 * n_embd, n_embd), nn.Dropout(dropout))

    def forward(self, x, residual=None):
        fc1, fc2 = self[0], self[2]
        _require_cuda(fc1.weight, "MLP weights")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        xb = x2 if x2.dtype == bf16 else _CastBf16.apply(x2.float())
        w1, w2 = flat_shadow(fc1.weight), flat_shadow(fc2.weight)
        if w1 is None or w2 is None:
            w1, w2 = self._get_shadow("mlp", (fc1.weight, fc2.weight),
                                      lambda: (ops.cast_bf16(fc1.weight), ops.cast_bf16(fc2.weight)))
        r2 = None if residual is None else residual.reshape(-1, shp[-1])
        drop = self.training and self[3].p > 0.0
        out = Fn.MlpGeluFn.apply(xb.contiguous(), None if drop else r2, w1, fc1.bias, w2, fc2.bias, fc1.weight,
                                 fc2.weight)
        if drop:  # nn.Dropout is the last child of the Sequential (:147); the residual add follows it (:152)
            out = Fn.DropoutFn.apply(out, r2, float(self[3].p))
        return out.view(shp[:-1] + (out.shape[-1],))


class CausalSelfAttention(nn.Module, _ShadowMixin):
    def __init__(self, n_embd, n_head, dropout, block_size, n_kv_head: int | None = None, use_sdpa: bool = False,
                 use_rope: bool = False):
        super().__init__()
        assert n_embd % n_head == 0
        self.n_head = n_head
        self.n_kv_head = n_kv_head if (n_kv_head is not None and n_kv_head > 0 and n_kv_head <= n_head) else None
        self.use_sdpa = bool(use_sdpa)
        head_dim = n_embd // n_head
        kv_dim = (self.n_kv_head * head_dim) if self.n_kv_head is not None else n_embd
        self.key = nn.Linear(n_embd, kv_dim)
        self.query = nn.Linear(n_embd, n_embd)
        self.value = nn.Linear(n_embd, kv_dim)
        self.proj = nn.Linear(n_embd, n_embd)
        self.dropout = nn.Dropout(dropout)
        self.register_buffer("mask", torch.tril(torch.ones(block_size, block_size)).unsqueeze(0).unsqueeze(0))
        self.rotary_emb = RotaryEmbedding(dim=head_dim, max_position_embeddings=block_size) if use_rope else None
        self.last_attn = None

    def packed_param_groups(self):
        """Parameters evaluated as one packed operand, in packed row order (trainer._packed_order)."""
        return [(self.query.weight, self.key.weight, self.value.weight),
                (self.query.bias, self.key.bias, self.value.bias)]

    def _shadows(self):
        q, k, v = self.query, self.key, self.value
        ws = [flat_shadow(m.weight) for m in (q, k, v)]
        wp = flat_shadow(self.proj.weight)
        bs = [q.bias.data, k.bias.data, v.bias.data]
        if wp is not None and all(w is not None for w in ws) and _adjacent(ws) and _adjacent(bs):
            # the three slices of the flat buffers already ARE the packed operands
            n, d = sum(w.shape[0] for w in ws), ws[0].shape[1]
            return torch.as_strided(ws[0], (n, d), (d, 1)), torch.as_strided(bs[0], (n,), (1,)), wp

        def build():
            d = q.weight.shape[1]
            nq, nk = q.weight.shape[0], k.weight.shape[0]
            w = torch.empty((nq + 2 * nk, d), dtype=bf16, device=q.weight.device)
            ops.cast_bf16(q.weight, out=w[:nq])
            ops.cast_bf16(k.weight, out=w[nq:nq + nk])
            ops.cast_bf16(v.weight, out=w[nq + nk:])
            b = torch.cat((q.bias, k.bias, v.bias)).float().contiguous()
            return w, b, ops.cast_bf16(self.proj.weight)
        return self._get_shadow("attn", (q.weight, k.weight, v.weight, q.bias, k.bias, v.bias, self.proj.weight), build)

    def forward(self, x, attn_mask=None, residual=None):
        if not self.query.weight.is_cuda:  # host-resident module called directly (tests/test_attention_dropout.py:33-59)
            return _host_call(self, "forward", (x,), dict(attn_mask=attn_mask, residual=residual), x.device)
        B, T, Cdim = x.size()
        H = self.n_head
        hd = Cdim // H
        Hk = self.n_kv_head if self.n_kv_head is not None else H
        if H % Hk != 0:
            raise ValueError("n_head must be divisible by n_kv_head for GQA")
        if isinstance(attn_mask, torch.Tensor):  # the reference's calling convention: a boolean mask tensor
            attn_mask = mask_spec_from_bool(attn_mask.to(x.device), B, T)
        spec: MaskSpec = attn_mask if attn_mask is not None else MaskSpec(None, 0)
        x2 = x.reshape(B * T, Cdim)
        xb = x2 if x2.dtype == bf16 else _CastBf16.apply(x2.float())
        w_qkv, b_qkv, w_proj = self._shadows()
        nq, nk = self.query.weight.shape[0], self.key.weight.shape[0]
        qkv = Fn.PackedLinearFn.apply(xb.contiguous(), w_qkv, b_qkv, None, ((0, nq), (nq, nk), (nq + nk, nk)), False,
                                      self.query.weight, self.key.weight, self.value.weight,
                                      self.query.bias, self.key.bias, self.value.bias)
        rope = self.rotary_emb.half_tables(T, x.device) if self.rotary_emb is not None else None
        attn_p = float(self.dropout.p) if self.training else 0.0  # dropout on the probabilities (:104,129)
        y = Fn.AttentionFn.apply(qkv, spec.seg_start, rope, B, T, H, Hk, hd, int(spec.window or 0), attn_p,
                                 (self.query.bias, self.key.bias, self.value.bias))
        sink = self.__dict__.get("_kv_sink")
        if sink is not None:  # prefill of the incremental-decode cache: the (rotated) key / value column blocks
            q3 = qkv.view(B, T, -1)
            sink[0][:, :T].copy_(q3[:, :, nq:nq + nk])
            sink[1][:, :T].copy_(q3[:, :, nq + nk:])
        if not self.use_sdpa:
            # the reference's manual branch keeps the (pre-dropout) probabilities for introspection (:128)
            with torch.no_grad():
                self.last_attn = ops.attn_probs(qkv.detach(), spec.seg_start, B, T, H, Hk, hd, int(spec.window or 0))
        r2 = None if residual is None else residual.reshape(B * T, Cdim)
        out = Fn.PackedLinearFn.apply(y, w_proj, self.proj.bias.detach(), r2, ((0, Cdim),), True, self.proj.weight,
                                      self.proj.bias)
        return out.view(B, T, Cdim)


    def decode(self, x, k_cache, v_cache, lo, t_dev, rope_row=None, window=0, residual=None):
        """One new position per sequence (x: (B, 1, d) bf16) against the K/V cache; appends the position held in the
        device scalar t_dev.  rope_row = (cos, sin) rows [1, hd/2] of that position."""
        B, _, Cdim = x.size()
        H = self.n_head
        hd = Cdim // H
        Hk = self.n_kv_head if self.n_kv_head is not None else H
        if H % Hk != 0:
            raise ValueError("n_head must be divisible by n_kv_head for GQA")
        xb = x.reshape(B, Cdim)
        w_qkv, b_qkv, w_proj = self._shadows()
        nq, nk = self.query.weight.shape[0], self.key.weight.shape[0]
        qkv = Fn.PackedLinearFn.apply(xb.contiguous(), w_qkv, b_qkv, None, ((0, nq), (nq, nk), (nq + nk, nk)), False,
                                      self.query.weight, self.key.weight, self.value.weight,
                                      self.query.bias, self.key.bias, self.value.bias)
        if self.rotary_emb is not None:
            ops.rope_qk(qkv, rope_row[0], rope_row[1], B, 1, H, Hk, hd)
        y = ops.attn_decode(qkv, k_cache, v_cache, lo, 0, H, Hk, hd, window=window, t_dev=t_dev)
        r2 = None if residual is None else residual.reshape(B, Cdim)
        out = Fn.PackedLinearFn.apply(y, w_proj, self.proj.bias.detach(), r2, ((0, Cdim),), True, self.proj.weight,
                                      self.proj.bias)
        return out.view(B, 1, Cdim)


class Block(nn.Module):
    def __init__(self, n_embd, n_head, dropout, block_size, n_kv_head: int | None = None, use_sdpa: bool = False,
                 use_swiglu: bool = False, use_rope: bool = False):
        super().__init__()
        self.ln1 = LayerNorm(n_embd)
        self.attn = CausalSelfAttention(n_embd, n_head, dropout, block_size, n_kv_head=n_kv_head, use_sdpa=use_sdpa,
                                        use_rope=use_rope)
        self.ln2 = LayerNorm(n_embd)
        self.mlp = SwiGLU(n_embd, dropout) if use_swiglu else GeluMLP(n_embd, dropout)

    def residual_bias(self):
        """The bias added right before this block's output joins the residual stream (fc2 of a GELU MLP), when that
        add is fused into the GEMM and nothing (dropout, hooks) sits behind it: the next LayerNorm's backward then adds
        its column sums straight into this bias' gradient slot.  None otherwise."""
        mlp = self.mlp
        if not isinstance(mlp, GeluMLP) or _has_hooks(mlp) or (self.training and mlp[3].p > 0.0):
            return None
        return mlp[2].bias

    def forward(self, x, attn_mask=None, prev_bias=None):
        B, T, d = x.shape
        x2 = x.reshape(B * T, d)
        if x2.dtype != f32:
            x2 = x2.float()
        # x = x + attn(ln1(x)): LN returns the residual stream so that backward fuses the two gradient paths;
        # the residual add itself happens in the epilogue of the proj / fc2 GEMM.
        if _has_hooks(self.ln1):
            h = _CastBf16.apply(self.ln1(x2))
        else:
            x2, h = self.ln1.fused(x2.contiguous(), prev_bias)
        if _has_hooks(self.attn):
            x2 = x2 + self.attn(h.view(B, T, d), attn_mask=attn_mask).reshape(B * T, d)
            proj_bias = None
        else:
            x2 = self.attn(h.view(B, T, d), attn_mask=attn_mask, residual=x2).reshape(B * T, d)
            proj_bias = self.attn.proj.bias
        if _has_hooks(self.ln2):
            h = _CastBf16.apply(self.ln2(x2))
        else:
            x2, h = self.ln2.fused(x2.contiguous(), proj_bias)
        if _has_hooks(self.mlp):
            x2 = x2 + self.mlp(h.view(B, T, d)).reshape(B * T, d)
        else:
            x2 = self.mlp(h.view(B, T, d), residual=x2).reshape(B * T, d)
        return x2.view(B, T, d)


def _block_decode(blk, x, k_cache, v_cache, lo, t_dev, rope_row, window):
    """Block.forward for one new position per sequence, attention against the K/V cache (eval mode, no hooks)."""
    B, _, d = x.shape
    x2 = x.reshape(B, d)
    x2, h = blk.ln1.fused(x2.contiguous())
    x2 = blk.attn.decode(h.view(B, 1, d), k_cache, v_cache, lo, t_dev, rope_row, window, residual=x2).reshape(B, d)
    x2, h = blk.ln2.fused(x2.contiguous())
    x2 = blk.mlp(h.view(B, 1, d), residual=x2).reshape(B, d)
    return x2.view(B, 1, d)


class DecodeState:
    """K/V caches of an incremental decode: per layer bf16 [B, max_len, Hk*hd], the number of cached positions and
    the first visible position of every sequence (start of its current <SEP> segment)."""

    def __init__(self, model, batch: int, max_len: int):
        dev = model.tok_emb.weight.device
        hd = model.n_embd // model.n_head
        hk = model.n_kv_head if model.n_kv_head is not None else model.n_head
        self.max_len = int(max_len)
        self.batch = int(batch)
        self.k = [torch.empty((batch, self.max_len, hk * hd), dtype=bf16, device=dev) for _ in range(model.n_layer)]
        self.v = [torch.empty((batch, self.max_len, hk * hd), dtype=bf16, device=dev) for _ in range(model.n_layer)]
        self.length = 0
        self.seg_lo = torch.zeros((batch,), dtype=torch.int32, device=dev)
        self.window = 0
        # device-resident step state: every kernel of a decode step reads the position from t_dev, so the step can be
        # captured ONCE into a CUDA graph and replayed for every position (a step is ~13 launches per layer and
        # entirely launch-bound when issued from Python)
        self.t_dev = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.tok = torch.zeros((batch, 1), dtype=torch.int64, device=dev)
        self.logits = None
        self.graph = None
        self.use_graph = True


class _OffsetMLP(nn.Sequential, _ShadowMixin):
    """nn.Sequential(Linear(d,d), GELU(), Linear(d,d)) per offset (model_tiny_gpt.py:235-239)."""

    def forward(self, x):
        fc1, fc2 = self[0], self[2]
        _require_cuda(fc1.weight, "offset head weights")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        xb = x2 if x2.dtype == bf16 else _CastBf16.apply(x2.float())
        w1, w2 = flat_shadow(fc1.weight), flat_shadow(fc2.weight)
        if w1 is None or w2 is None:
            w1, w2 = self._get_shadow("off", (fc1.weight, fc2.weight),
                                      lambda: (ops.cast_bf16(fc1.weight), ops.cast_bf16(fc2.weight)))
        out = Fn.OffsetHeadFn.apply(xb.contiguous(), w1, fc1.bias, w2, fc2.bias, fc1.weight, fc2.weight)
        return out.view(shp)

    def shadows(self):
        """bf16 operands (w1, w2): views of the trainer's flat shadow buffer, else cached casts of the masters."""
        fc1, fc2 = self[0], self[2]
        _require_cuda(fc1.weight, "offset head weights")
        w1, w2 = flat_shadow(fc1.weight), flat_shadow(fc2.weight)
        if w1 is None or w2 is None:
            w1, w2 = self._get_shadow("off", (fc1.weight, fc2.weight),
                                      lambda: (ops.cast_bf16(fc1.weight), ops.cast_bf16(fc2.weight)))
        return w1, w2


# ----------------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------------
class TinyGPT(nn.Module):
    def __init__(
        self,
        vocab_size,
        block_size,
        n_layer=3,
        n_head=4,
        n_embd=256,
        dropout=0.1,
        use_checkpoint=False,
        label_smoothing: float = 0.0,
        sep_id: int | None = 3,
        tie_embeddings: bool = True,
        n_kv_head: int | None = None,
        use_sdpa: bool = False,
        loss_weights: list[float] | None = None,
        termination_aux: bool = False,
        termination_n_classes: int = 5,
        multi_offset_targets: list[int] | None = None,
        use_swiglu: bool = False,
        use_rope: bool = False,
        use_shape_guidance: bool = False,
    ):
        super().__init__()
        self.block_size = block_size
        self.vocab_size = vocab_size
        self.n_layer = n_layer
        self.n_head = n_head
        self.n_embd = n_embd
        self.dropout_p = float(dropout)
        self.use_checkpoint = use_checkpoint  # accepted for config parity; activations fit in 180 GB, no recompute
        self.label_smoothing = float(label_smoothing)
        self.sep_id = sep_id
        self.tie_embeddings = bool(tie_embeddings)
        self.n_kv_head = n_kv_head if (n_kv_head is not None and n_kv_head > 0) else None
        self.use_sdpa = bool(use_sdpa)
        self.termination_aux = bool(termination_aux)
        self.termination_n_classes = int(termination_n_classes)
        self.use_swiglu = bool(use_swiglu)
        self.use_rope = bool(use_rope)
        self.use_shape_guidance = bool(use_shape_guidance)

        self.tok_emb = Embedding(vocab_size, n_embd)
        self.pos_emb = Embedding(block_size, n_embd) if not self.use_rope else None
        self.drop = nn.Dropout(dropout)
        self.blocks = nn.ModuleList([
            Block(n_embd, n_head, dropout, block_size, n_kv_head=self.n_kv_head, use_sdpa=self.use_sdpa,
                  use_swiglu=self.use_swiglu, use_rope=self.use_rope)
            for _ in range(n_layer)
        ])
        self.ln_f = LayerNorm(n_embd)
        self.head = SkinnyLinear(n_embd, vocab_size, bias=False)
        if self.tie_embeddings:
            self.head.weight = self.tok_emb.weight
        self.termination_head = SkinnyLinear(n_embd, self.termination_n_classes) if self.termination_aux else None

        if self.use_shape_guidance:
            self.shape_proj = nn.Linear(3, n_embd)
            nn.init.zeros_(self.shape_proj.weight)
            nn.init.zeros_(self.shape_proj.bias)

        self.multi_offset_targets = sorted(list(set([int(t) for t in multi_offset_targets]))) if multi_offset_targets else []
        self.offset_projs = nn.ModuleDict()
        for offset in self.multi_offset_targets:
            mlp = _OffsetMLP(nn.Linear(n_embd, n_embd), nn.GELU(), nn.Linear(n_embd, n_embd))
            nn.init.eye_(mlp[0].weight)
            nn.init.zeros_(mlp[0].bias)
            nn.init.eye_(mlp[2].weight)
            nn.init.zeros_(mlp[2].bias)
            self.offset_projs[str(offset)] = mlp

        if loss_weights is not None:
            self.register_buffer("loss_weights", torch.tensor(loss_weights, dtype=torch.float32))
        else:
            self.register_buffer("loss_weights", torch.ones(vocab_size, dtype=torch.float32))
        self._lw_cache = None

    def packed_param_groups(self):
        """The first linears of the offset MLPs all read the final hidden state: made neighbours in the trainer's flat
        buffers (trainer._packed_order) they are one [n_off*d, d] operand (functional.HeadsFn)."""
        mlps = [self.offset_projs[str(o)] for o in self.multi_offset_targets]
        if len(mlps) < 2:
            return []
        return [tuple(m[0].weight for m in mlps), tuple(m[0].bias for m in mlps)]

    # ------------------------------------------------------------------ reference API
    def to_dict(self) -> dict:
        return {
            "vocab_size": int(self.vocab_size),
            "block_size": int(self.block_size),
            "n_layer": int(self.n_layer),
            "n_head": int(self.n_head),
            "n_embd": int(self.n_embd),
            "dropout": float(self.dropout_p),
            "sep_mask_enabled": self.sep_id is not None,
            "tie_embeddings": bool(self.tie_embeddings),
            "n_kv_head": self.n_kv_head,
            "use_sdpa": bool(self.use_sdpa),
            "termination_aux": bool(self.termination_aux),
            "termination_n_classes": int(self.termination_n_classes),
            "multi_offset_targets": self.multi_offset_targets,
            "use_swiglu": bool(self.use_swiglu),
            "use_rope": bool(self.use_rope),
            "use_shape_guidance": bool(self.use_shape_guidance),
        }

    def build_attention_mask(self, idx: torch.Tensor, attention_window: int | None = None) -> torch.Tensor | None:
        """The boolean (B,1,T,T) mask of the reference (:273-295), for introspection only: the kernels use
        mask_spec() and never materialise it.  Index arithmetic only, any device."""
        _, length = idx.shape
        if attention_window is not None and int(attention_window) < 1:
            raise ValueError("attention_window must be at least 1")
        if self.sep_id is None and attention_window is None:
            return None
        pos = torch.arange(length, device=idx.device)
        dist = pos.unsqueeze(1) - pos.unsqueeze(0)
        allowed = dist >= 0
        if attention_window is not None:
            allowed = allowed & (dist < int(attention_window))
        allowed = allowed.unsqueeze(0).unsqueeze(0)
        if self.sep_id is not None:
            seg = (ops.segment_ids(idx.contiguous(), int(self.sep_id)).long() if idx.is_cuda
                   else torch.cumsum(idx == int(self.sep_id), dim=1))
            allowed = allowed & (seg.unsqueeze(-1) == seg.unsqueeze(-2)).unsqueeze(1)
        return allowed

    def mask_spec(self, idx: torch.Tensor, attention_window: int | None = None) -> MaskSpec:
        if attention_window is not None and int(attention_window) < 1:
            raise ValueError("attention_window must be at least 1")
        seg_start = ops.segment_starts(idx, int(self.sep_id)) if self.sep_id is not None else None
        return MaskSpec(seg_start, int(attention_window) if attention_window is not None else 0)

    # ------------------------------------------------------------------ internals
    def _prep_idx(self, idx):
        dev = self.tok_emb.weight.device
        _require_cuda(self.tok_emb.weight, "TinyGPT parameters")
        if idx.device != dev:
            idx = idx.to(dev)
        if idx.dtype != torch.int64:
            idx = idx.long()
        B, T = idx.shape
        if not self.use_rope and T > self.block_size:
            raise IndexError(f"sequence length {T} exceeds block_size {self.block_size}")
        return idx.contiguous()

    def _embed(self, idx, shape_embeddings):
        B, T = idx.shape
        if _has_hooks(self.tok_emb) or (self.pos_emb is not None and _has_hooks(self.pos_emb)):
            x = self.tok_emb(idx)
            if self.pos_emb is not None:
                x = x + self.pos_emb(torch.arange(0, T, device=idx.device).unsqueeze(0))
        else:
            x = Fn.EmbedFn.apply(idx, self.tok_emb.weight, None if self.pos_emb is None else self.pos_emb.weight)
        if shape_embeddings is not None and self.use_shape_guidance:  # x + shape_proj(shape_embeddings) (:310-311)
            if self.n_embd % 4 != 0:
                raise _lib.CgptError("use_shape_guidance needs n_embd % 4 == 0")
            s2 = shape_embeddings.to(device=x.device, dtype=f32).reshape(B * T, 3).contiguous()
            x = Fn.ShapeProjFn.apply(x.reshape(B * T, -1).contiguous(), s2, self.shape_proj.weight,
                                     self.shape_proj.bias).view(B, T, -1)
        if self.training and self.drop.p > 0.0:  # self.drop(x) (:312)
            x = Fn.DropoutFn.apply(x, None, float(self.drop.p))
        return x

    def class_weights(self):
        """loss_weights, or None when they are all 1 (the reference checks this with a host sync on every
        forward, :341; here once per buffer version)."""
        lw = self.loss_weights
        key = (lw.data_ptr(), lw._version)
        if self._lw_cache is None or self._lw_cache[0] != key:
            self._lw_cache = (key, bool(torch.all(lw == 1.0).item()))
        return None if self._lw_cache[1] else lw

    def _heads_fused_ok(self, M):
        mods = [self.head] + ([self.termination_head] if self.termination_head is not None else [])
        for mlp in self.offset_projs.values():
            mods += [mlp, mlp[0], mlp[2]]
        return (M >= Fn.TC_HEAD_MIN_ROWS and self.n_embd % 8 == 0 and self.vocab_size % 4 == 0
                and self.vocab_size <= 128 and self.head.bias is None
                and (self.termination_head is None or self.termination_n_classes <= 128)
                and not any(_has_hooks(m) for m in mods))

    def _heads(self, x, B, T):
        """x: (B,T,d) fp32 after ln_f -> logits, aux dict."""
        aux = {}
        if self._heads_fused_ok(B * T):
            # one autograd node for everything that reads the final hidden state (Fn.HeadsFn)
            x2 = x.reshape(B * T, -1).float().contiguous()
            th = self.termination_head
            args = []
            for offset in self.multi_offset_targets:
                mlp = self.offset_projs[str(offset)]
                w1, w2 = mlp.shadows()
                args += [w1, mlp[0].bias, w2, mlp[2].bias, mlp[0].weight, mlp[2].weight]
            outs = Fn.HeadsFn.apply(x2, self.head.weight, None if th is None else th.weight,
                                    None if th is None else th.bias, *args)
            logits = outs[0].view(B, T, -1)
            # which bf16 form of the logit gradient this node's backward GEMMs read (Fn.CrossEntropyFn writes it from
            # the same kernel as the fp32 gradient): 2 = hi|lo|hi split (fp32-accurate main head), 1 = padded bf16 copy
            logits._cgpt_grad_form = 2
            k = 1
            if th is not None:
                aux["termination_logits"] = outs[1].view(B, T, -1)
                k = 2
            if len(self.offset_projs) > 0:
                aux["offset_logits"] = {o: outs[k + i].view(B, T, -1) for i, o in enumerate(self.multi_offset_targets)}
                for lg in aux["offset_logits"].values():
                    lg._cgpt_grad_form = 1
            return logits, aux
        logits = self.head(x)
        if self.termination_head is not None:
            aux["termination_logits"] = self.termination_head(x)
        if len(self.offset_projs) > 0:
            xb = _CastBf16.apply(x)
            offset_logits = {}
            for offset in self.multi_offset_targets:
                offset_logits[offset] = self.head(self.offset_projs[str(offset)](xb))
            aux["offset_logits"] = offset_logits
        return logits, aux

    def forward(self, idx, targets=None, return_aux: bool = False, shape_embeddings=None,
                attention_window: int | None = None):
        in_dev = idx.device
        if not self.tok_emb.weight.is_cuda:  # host-resident model (the reference's inference callers): stage it
            return _host_call(self, "forward", (idx,), dict(targets=targets, return_aux=return_aux,
                                                             shape_embeddings=shape_embeddings,
                                                             attention_window=attention_window), in_dev)
        idx = self._prep_idx(idx)
        B, T = idx.shape
        Fn.reset_side_channel()
        x = self._embed(idx, shape_embeddings)
        spec = self.mask_spec(idx, attention_window)
        prev_bias = None  # bias of the residual linear in front of the next LayerNorm (Block.residual_bias)
        for blk in self.blocks:
            x = blk(x, attn_mask=spec, prev_bias=prev_bias)
            prev_bias = None if _has_hooks(blk) else blk.residual_bias()
        x = self.ln_f(x, colsum_target=prev_bias)
        logits, aux = self._heads(x, B, T)
        loss = None
        if targets is not None:
            tg = targets.to(idx.device).long().contiguous()
            loss, _ = Fn.CrossEntropyFn.apply(logits.view(B * T, -1), tg, None, self.class_weights(), B, T, 0,
                                              self.label_smoothing, 0, False, getattr(logits, "_cgpt_grad_form", 0))
        if in_dev != idx.device:  # callers that keep their tensors on the host get host results back
            logits = logits.to(in_dev)
        if return_aux:
            return logits, loss, aux
        return logits, loss

    @torch.no_grad()
    def next_token_logits(self, idx, attention_window: int | None = None):
        """Last-position logits (B, V) = forward(idx)[0][:, -1] — what the reference's sampling loops read after a
        full forward (src/codonlm/generate.py:14-27 `_next_token_logits`, scripts/query_model.py `next_token`),
        for a whole batch of contexts: the final LayerNorm and the fp32 head run on B rows instead of B·T."""
        in_dev = idx.device
        if not self.tok_emb.weight.is_cuda:
            return _host_call(self, "next_token_logits", (idx,), dict(attention_window=attention_window), in_dev)
        idx = self._prep_idx(idx)
        Fn.reset_side_channel()
        x = self._embed(idx, None)
        spec = self.mask_spec(idx, attention_window)
        for blk in self.blocks:
            x = blk(x, attn_mask=spec)
        logits = self.head(self.ln_f(x[:, -1, :].contiguous()))
        return logits if in_dev == idx.device else logits.to(in_dev)

    @torch.no_grad()
    def prefill(self, idx, state: "DecodeState | None" = None, max_len: int | None = None,
                attention_window: int | None = None):
        """Full forward over the prompts idx (B, T) that also fills the K/V caches -> (last-position logits (B, V),
        state).  With decode_step() this is the cached form of the reference's sampling loops, which re-run the whole
        context for every generated token (generate.py:14-27, query_model.py:186-213): same logits, O(T) per token."""
        idx = self._prep_idx(idx)
        B, T = idx.shape
        if state is None:
            state = DecodeState(self, B, max_len or self.block_size)
        if T > state.max_len or B != state.batch:
            raise ValueError(f"prompt of {B}x{T} tokens does not fit a decode state of {state.batch}x{state.max_len}")
        Fn.reset_side_channel()
        x = self._embed(idx, None)
        spec = self.mask_spec(idx, attention_window)
        for layer, blk in enumerate(self.blocks):
            blk.attn.__dict__["_kv_sink"] = (state.k[layer], state.v[layer])
            try:
                x = blk(x, attn_mask=spec)
            finally:
                blk.attn.__dict__["_kv_sink"] = None
        state.length = T
        state.t_dev.fill_(T)
        if int(spec.window or 0) != state.window:
            state.window, state.graph = int(spec.window or 0), None
        if spec.seg_start is not None:
            state.seg_lo.copy_(spec.seg_start[:, -1])
        else:
            state.seg_lo.zero_()
        return self.head(self.ln_f(x[:, -1, :].contiguous())), state

    def _decode_launch(self, state: "DecodeState"):
        """The kernels of one decode step; reads state.tok / state.t_dev, advances t_dev, returns the logits."""
        dev = self.tok_emb.weight.device
        tok, t_dev = state.tok, state.t_dev
        if self.sep_id is not None:  # a <SEP> opens a new segment at its own position (cumsum(idx == sep), :290)
            state.seg_lo.copy_(torch.where(tok.view(-1) == int(self.sep_id), t_dev.expand(state.batch), state.seg_lo))
        t_idx = t_dev.long()
        pos_row = None if self.pos_emb is None else self.pos_emb.weight.index_select(0, t_idx)
        rope_row = None
        if self.use_rope:
            cos, sin = self.blocks[0].attn.rotary_emb.half_tables(state.max_len, dev)
            rope_row = (cos.index_select(0, t_idx), sin.index_select(0, t_idx))
        x = Fn.EmbedFn.apply(tok, self.tok_emb.weight, pos_row)
        lo = state.seg_lo if self.sep_id is not None else None
        for layer, blk in enumerate(self.blocks):
            x = _block_decode(blk, x, state.k[layer], state.v[layer], lo, t_dev, rope_row, state.window)
        logits = self.head(self.ln_f(x[:, -1, :].contiguous()))
        t_dev.add_(1)
        return logits

    @torch.no_grad()
    def decode_step(self, tokens, state: "DecodeState"):
        """Append one token per sequence (tokens (B,), the token AT position state.length) and return the logits
        (B, V) for the following position — equal to forward(context + token)[0][:, -1].  The step is captured into
        a CUDA graph on its second call and replayed afterwards (state.use_graph = False keeps it eager)."""
        t = state.length
        if t >= state.max_len or (self.pos_emb is not None and t >= self.block_size):
            raise IndexError(f"decode position {t} exceeds the cache / block size; re-prefill a cropped context "
                             "(the reference crops to the last block_size tokens, generate.py:20-21)")
        state.tok.copy_(tokens.reshape(state.batch, 1), non_blocking=True)
        Fn.reset_side_channel()
        # a captured step holds pointers to the bf16 weight copies it was captured with: drop it when a master changed
        wkey = (_SHADOW_GEN[0],) + tuple(p._version for p in self.parameters())
        if state.graph is not None and state.__dict__.get("_wkey") != wkey:
            state.graph = None
        if state.graph is not None:
            state.graph.replay()
        elif state.use_graph and getattr(state, "_warm", False) and not self.training:
            # capture leaves the state untouched: run the step inside the capture, then replay it once for real
            graph = torch.cuda.CUDAGraph()
            snap_t, snap_lo = state.t_dev.clone(), state.seg_lo.clone()
            with torch.cuda.graph(graph):
                state.logits = self._decode_launch(state)
            state.t_dev.copy_(snap_t)
            state.seg_lo.copy_(snap_lo)
            state.graph = graph
            state._wkey = wkey
            graph.replay()
        else:
            state.logits = self._decode_launch(state)
            state._warm = True
        state.length = t + 1
        return state.logits

    def forward_hidden(self, idx, shape_embeddings=None, attention_window: int | None = None):
        final = None
        for _, hidden in self.iter_hidden_states(idx, shape_embeddings=shape_embeddings,
                                                 attention_window=attention_window):
            final = hidden
        if final is None:
            raise RuntimeError("hidden-state iterator produced no states")
        return final

    def iter_hidden_states(self, idx, shape_embeddings=None, attention_window: int | None = None):
        """Yield canonical causal states at embedding, block, and final-norm stages (:368-389)."""
        if not self.tok_emb.weight.is_cuda:  # host-resident model (scripts/extract_embeddings.py:273-274): stage it
            home = idx.device
            _staging_device()
            if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
                raise _lib.CgptError("training a host-resident model: move it to the GPU first (model.to('cuda'))")
            twin = _device_twin(self)
            dev = next(twin.parameters()).device
            with torch.no_grad():
                for stage, hidden in twin.iter_hidden_states(_to_device(idx, dev),
                                                             shape_embeddings=_to_device(shape_embeddings, dev),
                                                             attention_window=attention_window):
                    yield stage, hidden.to(home)
            return
        idx = self._prep_idx(idx)
        x = self._embed(idx, shape_embeddings)
        spec = self.mask_spec(idx, attention_window)
        yield 0, x
        for layer, blk in enumerate(self.blocks, start=1):
            x = blk(x, attn_mask=spec)
            yield layer, x
        x = self.ln_f(x)
        yield "final", x


class NoPropBlock(Block):
    """Block + a per-block denoising head (model_tiny_gpt.py:391-416): forward(h, noisy_targets, attn_mask) ->
    (x, pred_y) with x = block(h + noisy_targets) and pred_y = denoise_head(x).  Same submodule names, creation order
    (seeded init) and state_dict keys as the reference; the block itself runs on the same kernels as TinyGPT's."""

    def __init__(self, n_embd, n_head, dropout, block_size, n_kv_head: int | None = None, use_sdpa: bool = False):
        super().__init__(n_embd, n_head, dropout, block_size, n_kv_head=n_kv_head, use_sdpa=use_sdpa)
        self.denoise_head = Linear(n_embd, n_embd)

    def forward(self, h, noisy_targets=None, attn_mask=None):
        if noisy_targets is not None:  # x = h + noisy_targets (:407-408), through the residual-add kernel
            shp = h.shape
            h2 = h.reshape(-1, shp[-1]).float().contiguous()
            n2 = noisy_targets.to(h.device).reshape(-1, shp[-1]).float().contiguous()
            h = Fn.DropoutFn.apply(n2, h2, 0.0).view(shp)
        x = super().forward(h, attn_mask=attn_mask)
        return x, self.denoise_head(x)


class NoPropTinyGPT(nn.Module):
    """The reference's NoProp variant (model_tiny_gpt.py:418-459): every block also predicts the clean target
    embedding; forward(idx, target_embeddings) -> (logits, [pred_y per block]).  Only the module is provided — the
    NoProp training procedure (src/codonlm/noprop_task.py, train_noprop.py) drives it unchanged."""

    def __init__(self, vocab_size, block_size, n_layer=3, n_head=4, n_embd=256, dropout=0.1, sep_id: int | None = 3,
                 n_kv_head: int | None = None, use_sdpa: bool = False):
        super().__init__()
        self.block_size = block_size
        self.vocab_size = vocab_size
        self.n_embd = n_embd
        self.sep_id = sep_id
        self.tok_emb = Embedding(vocab_size, n_embd)
        self.pos_emb = Embedding(block_size, n_embd)
        self.drop = nn.Dropout(dropout)
        self.blocks = nn.ModuleList([
            NoPropBlock(n_embd, n_head, dropout, block_size, n_kv_head=n_kv_head, use_sdpa=use_sdpa)
            for _ in range(n_layer)
        ])
        self.ln_f = LayerNorm(n_embd)
        self.head = SkinnyLinear(n_embd, vocab_size, bias=False)
        self.head.weight = self.tok_emb.weight

    def forward(self, idx, target_embeddings=None):
        _require_cuda(self.tok_emb.weight, "NoPropTinyGPT parameters")
        dev = self.tok_emb.weight.device
        idx = idx.to(dev).long().contiguous()
        B, T = idx.shape
        if T > self.block_size:
            raise IndexError(f"sequence length {T} exceeds block_size {self.block_size}")
        Fn.reset_side_channel()
        h = Fn.EmbedFn.apply(idx, self.tok_emb.weight, self.pos_emb.weight)
        if self.training and self.drop.p > 0.0:
            h = Fn.DropoutFn.apply(h, None, float(self.drop.p))
        # causal & same-segment (:441-446) in interval form; without sep_id the reference passes no mask (pure causal)
        spec = MaskSpec(ops.segment_starts(idx, int(self.sep_id)), 0) if self.sep_id is not None else None
        preds = []
        for blk in self.blocks:
            h, pred_y = blk(h, noisy_targets=target_embeddings, attn_mask=spec)
            preds.append(pred_y)
        return self.head(self.ln_f(h)), preds


__all__ = ["TinyGPT", "NoPropBlock", "NoPropTinyGPT", "DecodeState"]
