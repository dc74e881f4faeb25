"""Deterministic synthetic weights / inputs for the YUME DiT (TEST INFRASTRUCTURE, like the rest of oracle/).

The reference zero-initialises head.head (wan23/modules/model.py:914), which would make every output 0, and ships
no checkpoints here; tests and bench therefore use seeded random weights with the reference's state-dict keys and
shapes. All values are rounded to bf16 so the fp32 oracle and the bf16 CUDA path hold identical
numbers. CPU `torch.Generator` streams are platform independent, so a (config, seed) pair names the tensors.
"""
from __future__ import annotations

from typing import Dict

import torch

Tensor = torch.Tensor

# real geometries (wan23/textimage2video.py:129-142, wan/image2video.py:140-153)
CFG_5B = dict(variant="5b", dim=3072, ffn_dim=14336, num_heads=24, num_layers=30, in_dim=48, out_dim=48,
              text_len=512, text_dim=4096, freq_dim=256, clip_dim=1280)
CFG_14B = dict(variant="14b", dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, in_dim=36, out_dim=16,
               text_len=512, text_dim=4096, freq_dim=256, clip_dim=1280)
# tiny geometries with the real head_dim (128) for golden vectors and CPU-sized parity tests
CFG_5B_TINY = dict(variant="5b", dim=256, ffn_dim=512, num_heads=2, num_layers=2, in_dim=48, out_dim=48,
                   text_len=32, text_dim=64, freq_dim=256, clip_dim=1280)
CFG_14B_TINY = dict(variant="14b", dim=256, ffn_dim=512, num_heads=2, num_layers=2, in_dim=36, out_dim=16,
                    text_len=32, text_dim=64, freq_dim=256, clip_dim=96)
# 8-head geometries (dim 1024): the smallest models whose heads divide over 2, 4 AND 8 Ulysses ranks
CFG_5B_H8 = dict(variant="5b", dim=1024, ffn_dim=2048, num_heads=8, num_layers=2, in_dim=48, out_dim=48,
                 text_len=32, text_dim=64, freq_dim=256, clip_dim=1280)
CFG_14B_H8 = dict(variant="14b", dim=1024, ffn_dim=2048, num_heads=8, num_layers=2, in_dim=36, out_dim=16,
                  text_len=32, text_dim=64, freq_dim=256, clip_dim=96)


def oracle_kwargs(cfg: dict) -> dict:
    return {k: cfg[k] for k in ("variant", "dim", "ffn_dim", "num_heads", "num_layers", "in_dim", "out_dim",
                                "text_len", "freq_dim")}


def param_shapes(cfg: dict, num_layers: int | None = None) -> Dict[str, tuple]:
    """Every parameter of the reference WanModel (+ the FramePack embedders) in a fixed order."""
    C, Fd, cin, cout = cfg["dim"], cfg["ffn_dim"], cfg["in_dim"], cfg["out_dim"]
    nl = cfg["num_layers"] if num_layers is None else num_layers
    s: Dict[str, tuple] = {}
    for name, k in (("patch_embedding", 2), ("patch_embedding_2x", 4), ("patch_embedding_4x", 8),
                    ("patch_embedding_8x", 16), ("patch_embedding_16x", 32)):
        s[f"{name}.weight"] = (C, cin, 1, k, k)
        s[f"{name}.bias"] = (C,)
    s["patch_embedding_2x_f.weight"] = (cin, cin, 1, 4, 4)
    s["patch_embedding_2x_f.bias"] = (cin,)
    s["text_embedding.0.weight"], s["text_embedding.0.bias"] = (C, cfg["text_dim"]), (C,)
    s["text_embedding.2.weight"], s["text_embedding.2.bias"] = (C, C), (C,)
    s["time_embedding.0.weight"], s["time_embedding.0.bias"] = (C, cfg["freq_dim"]), (C,)
    s["time_embedding.2.weight"], s["time_embedding.2.bias"] = (C, C), (C,)
    s["time_projection.1.weight"], s["time_projection.1.bias"] = (6 * C, C), (6 * C,)
    if cfg["variant"] == "14b":
        cd = cfg["clip_dim"]
        s["img_emb.proj.0.weight"], s["img_emb.proj.0.bias"] = (cd,), (cd,)
        s["img_emb.proj.1.weight"], s["img_emb.proj.1.bias"] = (cd, cd), (cd,)
        s["img_emb.proj.3.weight"], s["img_emb.proj.3.bias"] = (C, cd), (C,)
        s["img_emb.proj.4.weight"], s["img_emb.proj.4.bias"] = (C,), (C,)
    for i in range(nl):
        p = f"blocks.{i}"
        for att in ("self_attn", "cross_attn"):
            projs = ["q", "k", "v", "o"] + (["k_img", "v_img"] if (att == "cross_attn" and cfg["variant"] == "14b") else [])
            for pr in projs:
                s[f"{p}.{att}.{pr}.weight"], s[f"{p}.{att}.{pr}.bias"] = (C, C), (C,)
            s[f"{p}.{att}.norm_q.weight"] = (C,)
            s[f"{p}.{att}.norm_k.weight"] = (C,)
            if att == "cross_attn" and cfg["variant"] == "14b":
                s[f"{p}.{att}.norm_k_img.weight"] = (C,)
        s[f"{p}.norm3.weight"], s[f"{p}.norm3.bias"] = (C,), (C,)
        s[f"{p}.ffn.0.weight"], s[f"{p}.ffn.0.bias"] = (Fd, C), (Fd,)
        s[f"{p}.ffn.2.weight"], s[f"{p}.ffn.2.bias"] = (C, Fd), (C,)
        s[f"{p}.modulation"] = (1, 6, C)
    s["head.head.weight"], s["head.head.bias"] = (4 * cout, C), (4 * cout,)
    s["head.modulation"] = (1, 2, C)
    return s


def make_state_dict(cfg: dict, seed: int, num_layers: int | None = None, device: str = "cpu") -> Dict[str, Tensor]:
    """Seeded fp32 tensors holding bf16-representable values. One generator per parameter (seeded by the global
    seed and the parameter's position) so a subset of layers reproduces the same leading tensors."""
    sd: Dict[str, Tensor] = {}
    for idx, (name, shape) in enumerate(param_shapes(cfg, num_layers).items()):
        g = torch.Generator(device="cpu").manual_seed(seed * 100003 + idx)
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith("modulation"):
            t = t / (cfg["dim"] ** 0.5) * 4.0  # larger than the reference init so gates/shifts matter numerically
        elif "norm" in name and name.endswith(".weight") or name in ("img_emb.proj.0.weight", "img_emb.proj.4.weight"):
            t = 1.0 + 0.1 * t
        elif name.endswith(".bias"):
            t = 0.02 * t
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t / (fan_in ** 0.5)
        sd[name] = t.to(torch.bfloat16).to(torch.float32).to(device)
    return sd


def make_inputs(cfg: dict, seed: int, frames: int, height: int, width: int, ctx_len: int, with_y: bool = False):
    """Gaussian latent [in_dim(-y), F, H, W], text context [ctx_len, text_dim] (bf16-representable), and for the 14B
    tree the conditioning latent y and CLIP features [1, 257, clip_dim]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cin = cfg["in_dim"]
    ycin = 20 if (cfg["variant"] == "14b") else 0
    x = torch.randn(cin - ycin, frames, height, width, generator=g)
    ctx = torch.randn(ctx_len, cfg["text_dim"], generator=g).to(torch.bfloat16).float()
    out = dict(x=x, context=ctx)
    if cfg["variant"] == "14b":
        out["y"] = torch.randn(ycin, frames, height, width, generator=g)
        out["clip_fea"] = torch.randn(1, 257, cfg["clip_dim"], generator=g).to(torch.bfloat16).float()
    return out
