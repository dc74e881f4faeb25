"""The finer drop-in seams of SURVEY.md §8(b), below `WanModel.forward`: the module-level `flash_attention` function and the
`WanAttentionBlock.forward` / `WanSelfAttention.forward` methods, each with the reference's signature, argument meaning and
error behaviour, each running on the kernels of libyume_b200.so.

  flash_attention            wan23/modules/attention.py:24-130 (identical in wan/modules/attention.py)
  WanAttentionBlock.forward  wan23/modules/model.py:272-316 (5B)   wan/modules/model.py:444-496 (14B)
  WanSelfAttention.forward   wan23/modules/model.py:178-207 (5B)   wan/modules/model.py:289-320 (14B)

`yume_b200.install(model)` re-binds the whole forward (the fast path: one engine call per denoise step). These seams exist for
callers that drive blocks or attention themselves (the reference's own sequence-parallel patch re-binds `self_attn.forward`,
wan23/textimage2video.py:190-194; `ulysses.distributed_attention` calls `flash_attention`): `install_seams(model)` binds them
on every block of a model that already carries an engine, `patch_flash_attention(module)` swaps the function a reference
`model.py` imported by name.
"""
from __future__ import annotations

import math
import types
from typing import Optional

import torch

from . import ops
from ._lib import YumeB200Error

__all__ = ["flash_attention", "patch_flash_attention", "install_seams", "block_forward_5b", "block_forward_14b",
           "self_attn_forward_5b", "self_attn_forward_14b"]


def flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
                    window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    """Drop-in for the reference `flash_attention` (attention.py:24-130): q [B, Lq, N, D], k / v [B, Lk, N, D] ->
    [B, Lq, N, D] in q's dtype; non-causal softmax(q k^T * softmax_scale) v, keys >= k_lens[b] dropped; inputs that are not
    fp16/bf16 are cast to `dtype` as the reference's `half()` does. Supported: what the Yume paths use — head_dim 128,
    Nk == Nq, no dropout, no causal mask, no window; anything else raises instead of computing something different."""
    half_dtypes = (torch.float16, torch.bfloat16)
    assert dtype in half_dtypes                                            # attention.py:53
    assert q.device.type == "cuda" and q.size(-1) <= 256                   # attention.py:54
    if q.size(-1) != 128 or v.size(-1) != 128:
        raise NotImplementedError("yume_b200 attention kernel: head_dim 128 only (both Yume models)")
    if k.size(2) != q.size(2):
        raise NotImplementedError("grouped-query attention (Nk != Nq) is not used by the Yume models")
    if dropout_p != 0.0 or causal or tuple(window_size) != (-1, -1):
        raise NotImplementedError("dropout / causal / sliding-window attention are not used by the Yume models")
    b, lq, lk, n, out_dtype = q.size(0), q.size(1), k.size(1), q.size(2), q.dtype
    if q_lens is not None and any(int(x) != lq for x in q_lens):
        # the reference unflattens the packed result to [b, lq] (attention.py:130): ragged q_lens fail there as well
        raise RuntimeError("q_lens shorter than Lq cannot be unflattened to [B, Lq] (reference attention.py:130)")
    scale = (1.0 / math.sqrt(q.size(-1))) if softmax_scale is None else float(softmax_scale)
    if q_scale is not None:
        scale = scale * float(q_scale)                                     # q = q * q_scale (attention.py:71-72), folded
    out = torch.empty(b, lq, n, 128, device=q.device, dtype=torch.bfloat16)
    with torch.cuda.device(q.device):
        for i in range(b):
            kl = lk if k_lens is None else int(k_lens[i])
            if kl <= 0:
                raise YumeB200Error("k_lens must be positive")
            qi = q[i].to(torch.bfloat16).reshape(lq, n * 128)
            ki = k[i, :kl].to(torch.bfloat16).reshape(kl, n * 128)
            vi = v[i, :kl].to(torch.bfloat16).reshape(kl, n * 128)
            ops.attention(qi, ki, vi, out[i].view(lq, n * 128), n, scale=scale)
    return out.type(out_dtype)


def patch_flash_attention(*modules) -> None:
    """Point the `flash_attention` name of reference modules at the function above (model.py does
    `from .attention import flash_attention`, so both the attention module and every importer need the new binding)."""
    for m in modules:
        if hasattr(m, "flash_attention"):
            m.flash_attention = flash_attention


# ------------------------------------------------------------------------------------------------------------
# block / self-attention methods (bound by install_seams)
# ------------------------------------------------------------------------------------------------------------
def _eng(self):
    eng = getattr(self, "_yb_engine", None)
    if eng is None:
        raise YumeB200Error("yume_b200.install_seams(model) has not been called on this module's model")
    return eng, self._yb_index


def _no_mvdt(ids_keep, ids_restore):
    if ids_keep is not None or ids_restore is not None:
        raise NotImplementedError("ids_keep / ids_restore belong to the MVDT masked-training path; inference only")


def block_forward_5b(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens, ids_keep=None, ids_restore=None,
                     mask_token=None, flag=True):
    """WanAttentionBlock.forward of the 5B tree (wan23/modules/model.py:272-316). x [B, L, C], e [B, L, 6, C] fp32."""
    assert e.dtype == torch.float32                                        # model.py:294
    _no_mvdt(ids_keep, ids_restore)
    if context_lens is not None:
        raise NotImplementedError("context_lens is None on every Yume path (model.py:815)")
    eng, i = _eng(self)
    outs = [eng.block_forward(i, x[b], e[b], None if flag else tuple(int(v) for v in grid_sizes[b]), context[b],
                              freqs=freqs, packed=bool(flag), k_len=int(seq_lens[b]))
            for b in range(x.shape[0])]
    return torch.stack(outs)


def block_forward_14b(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens, rand_num_img=None, ids_keep=None,
                      ids_restore=None, mask_token=None, seq_lens1=None, cnt_blocks=None):
    """WanAttentionBlock.forward of the 14B tree (wan/modules/model.py:444-496). x [B, L, C], e [B, 6, C] fp32."""
    assert e.dtype == torch.float32                                        # wan/modules/model.py:468
    _no_mvdt(ids_keep, ids_restore)
    packed = not (rand_num_img is not None and rand_num_img < 0.4)         # rope_apply dispatch, wan/modules/model.py:40-41
    eng, i = _eng(self)
    outs = [eng.block_forward(i, x[b], e[b], None if packed else tuple(int(v) for v in grid_sizes[b]), context[b],
                              freqs=freqs, packed=packed, k_len=int(seq_lens[b]))
            for b in range(x.shape[0])]
    return torch.stack(outs)


def self_attn_forward_5b(self, x, seq_lens, grid_sizes, freqs, mask_token=None, ids_restore=None, ids_keep=None, flag=True):
    """WanSelfAttention.forward of the 5B tree (wan23/modules/model.py:178-207): [B, L, C] -> [B, L, C] (bf16, as the
    reference's autocast Linear returns)."""
    _no_mvdt(ids_keep, ids_restore)
    eng, i = _eng(self)
    return torch.stack([eng.self_attention_forward(i, x[b], None if flag else tuple(int(v) for v in grid_sizes[b]), freqs,
                                                   bool(flag), k_len=int(seq_lens[b])) for b in range(x.shape[0])])


def self_attn_forward_14b(self, x, seq_lens, grid_sizes, freqs, ids_restore=None, ids_keep=None, mask_token=None,
                          rand_num_img=None):
    """WanSelfAttention.forward of the 14B tree (wan/modules/model.py:289-320)."""
    _no_mvdt(ids_keep, ids_restore)
    packed = not (rand_num_img is not None and rand_num_img < 0.4)
    eng, i = _eng(self)
    return torch.stack([eng.self_attention_forward(i, x[b], None if packed else tuple(int(v) for v in grid_sizes[b]), freqs,
                                                   packed, k_len=int(seq_lens[b])) for b in range(x.shape[0])])


def install_seams(model: torch.nn.Module, which=("block", "self_attn")) -> torch.nn.Module:
    """Bind the block / self-attention forwards on every block of a model that `yume_b200.install` has given an engine."""
    eng = getattr(model, "_yb_engine", None)
    if eng is None:
        raise YumeB200Error("call yume_b200.install(model) first")
    five = eng.variant == "5b"
    for i, blk in enumerate(model.blocks):
        for mod in (blk, blk.self_attn):
            mod._yb_engine, mod._yb_index = eng, i
        if "block" in which:
            blk.forward = types.MethodType(block_forward_5b if five else block_forward_14b, blk)
        if "self_attn" in which:
            blk.self_attn.forward = types.MethodType(self_attn_forward_5b if five else self_attn_forward_14b, blk.self_attn)
    return model
