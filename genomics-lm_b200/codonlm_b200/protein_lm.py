"""`ProteinConditionalTransformer` of the reference (src/protein_lm/models.py:5-59) on the same sm_100a kernels
(SURVEY §8f-4): a causal GPT over protein tokens built from `nn.TransformerEncoderLayer` blocks — POST-norm
(`x = norm1(x + attn(x)); x = norm2(x + ff(x))`), packed `in_proj` q|k|v, exact-erf GELU, learned positions, untied
bias-free output head.

The parameters live in real `nn.TransformerEncoderLayer` / `nn.Embedding` / `nn.LayerNorm` / `nn.Linear` modules created
in the reference's order, so `torch.manual_seed(s)` gives the reference's initial weights and the state_dict keys
(`transformer_blocks.<i>.self_attn.in_proj_weight`, `...linear1.weight`, `...norm1.weight`, ...) are the reference's:
its checkpoints load strictly.  Only the arithmetic differs: packed-QKV tcgen05 GEMM, causal flash attention, GEMM
epilogues with bias / GELU / fp32 residual, LayerNorm kernels, fp32 head.

Not provided: `ProteinClassifier` (models.py:61-117) attends BIDIRECTIONALLY with a key-padding mask — not a causal
interval mask, which is all the attention kernels implement; dropout INSIDE the feed-forward (between GELU and linear2)
in training mode (the two GEMMs are fused around the activation): a model with dropout > 0 trains only in eval-free
paths, i.e. `forward` raises in training mode when config.dropout > 0.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from . import functional as Fn
from .model_tiny_gpt import _CastBf16, _host_call, _ShadowMixin, _require_cuda  # noqa: F401


class ProteinConditionalTransformer(nn.Module, _ShadowMixin):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.token_embedding = nn.Embedding(config.vocab_size, config.n_embd)
        self.position_embedding = nn.Embedding(config.block_size, config.n_embd)
        self.dropout = nn.Dropout(config.dropout)
        self.transformer_blocks = nn.ModuleList([
            nn.TransformerEncoderLayer(d_model=config.n_embd, nhead=config.n_head, dim_feedforward=4 * config.n_embd,
                                       dropout=config.dropout, batch_first=True, activation="gelu")
            for _ in range(config.n_layer)
        ])
        self.layer_norm = nn.LayerNorm(config.n_embd)
        self.output_head = nn.Linear(config.n_embd, config.vocab_size, bias=False)

    def _block_shadows(self, i, blk):
        sa = blk.self_attn
        params = (sa.in_proj_weight, sa.out_proj.weight, blk.linear1.weight, blk.linear2.weight)
        return self._get_shadow(f"blk{i}", params, lambda: tuple(ops.cast_bf16(p) for p in params))

    def forward(self, input_ids: torch.LongTensor) -> torch.Tensor:
        w = self.token_embedding.weight
        if not w.is_cuda:  # host-resident model: stage it (inference), as TinyGPT does
            return _host_call(self, "forward", (input_ids,), {}, input_ids.device)
        cfg = self.config
        if self.training and float(cfg.dropout) > 0.0:
            raise _lib.CgptError("ProteinConditionalTransformer: training with dropout > 0 is not implemented on the "
                                 "fused feed-forward (dropout sits between GELU and linear2); use dropout 0 or eval()")
        idx = input_ids.to(w.device).long().contiguous()
        B, T = idx.shape
        if T > cfg.block_size:
            raise IndexError(f"sequence length {T} exceeds block_size {cfg.block_size}")
        d, H = cfg.n_embd, cfg.n_head
        hd = d // H
        if d % H != 0 or hd % 16 != 0 or hd > 128:
            raise _lib.CgptError(f"head_dim {hd} must be a multiple of 16 and <= 128")
        Fn.reset_side_channel()
        x = Fn.EmbedFn.apply(idx, w, self.position_embedding.weight).reshape(B * T, d)  # fp32 residual stream
        xb = _CastBf16.apply(x)
        for i, blk in enumerate(self.transformer_blocks):
            sa = blk.self_attn
            w_in, w_out, w1, w2 = self._block_shadows(i, blk)
            # the three d-row slices of in_proj are views: their gradients accumulate into in_proj_weight.grad
            wq, wk, wv = sa.in_proj_weight[:d], sa.in_proj_weight[d:2 * d], sa.in_proj_weight[2 * d:]
            bq, bk, bv = sa.in_proj_bias[:d], sa.in_proj_bias[d:2 * d], sa.in_proj_bias[2 * d:]
            qkv = Fn.PackedLinearFn.apply(xb.contiguous(), w_in, sa.in_proj_bias.detach(), None,
                                          ((0, d), (d, d), (2 * d, d)), False, wq, wk, wv, bq, bk, bv)
            y = Fn.AttentionFn.apply(qkv, None, None, B, T, H, H, hd, 0, 0.0, None)  # causal (models.py:50)
            x = Fn.PackedLinearFn.apply(y, w_out, sa.out_proj.bias.detach(), x, ((0, d),), True, sa.out_proj.weight,
                                        sa.out_proj.bias)                                            # x + attn(x)
            _, xb, x = Fn.ResidualLayerNormFn.apply(x.contiguous(), blk.norm1.weight, blk.norm1.bias, True, None)
            x = Fn.MlpGeluFn.apply(xb.contiguous(), x.contiguous(), w1, blk.linear1.bias, w2, blk.linear2.bias,
                                   blk.linear1.weight, blk.linear2.weight)                           # x + ff(x)
            _, xb, x = Fn.ResidualLayerNormFn.apply(x.contiguous(), blk.norm2.weight, blk.norm2.bias, True, None)
        _, _, xf = Fn.ResidualLayerNormFn.apply(x.contiguous(), self.layer_norm.weight, self.layer_norm.bias, True, None)
        V = cfg.vocab_size
        if V > 128:
            raise _lib.CgptError(f"output head with {V} outputs > 128 is not supported by the fp32 head kernels")
        use_tc = B * T >= Fn.TC_HEAD_MIN_ROWS and d % 8 == 0 and V % 4 == 0
        logits = (Fn.SplitHeadFn if use_tc else Fn.SkinnyLinearFn).apply(xf.contiguous(), self.output_head.weight, None)
        return logits.view(B, T, V)


__all__ = ["ProteinConditionalTransformer"]
