"""Generate tests/golden/*.pt by running the REFERENCE's own WanModel code (imported from /root/reference) on CPU.

Run in the authoring container only (the reference tree does not exist on the GPU box):
    python tools/make_golden.py
Recipe: stub `diffusers` (absent here), register bare `wan*` packages so their heavy
__init__ files never run, load attention.py, replace `flash_attention` (which asserts CUDA and calls the
third-party flash_attn kernel) by an SDPA restatement of its contract, then load model.py by path.
Weights/inputs come from oracle/synth.py (seeded); only small outputs are stored.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from copy import deepcopy
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
from oracle import synth  # noqa: E402


def _sdpa_flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None,
                          causal=False, window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    """Same contract as reference flash_attention (attention.py:24-130) on CPU."""
    assert q_lens is None and softmax_scale is None and q_scale is None and not causal
    out_dtype = q.dtype
    b, lk = q.shape[0], k.shape[1]
    half = lambda x: x if x.dtype in (torch.float16, torch.bfloat16) else x.to(dtype)  # noqa: E731
    qh, kh, vh = half(q).transpose(1, 2), half(k).transpose(1, 2), half(v).transpose(1, 2)
    mask = None
    if k_lens is not None:
        mask = torch.zeros(b, 1, 1, lk, dtype=torch.bool)
        for i, n in enumerate(k_lens):
            mask[i, ..., :int(n)] = True
    return F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask).transpose(1, 2).contiguous().type(out_dtype)


def load_reference(tree: str):
    """tree: 'wan23' (5B) or 'wan' (14B). Returns the loaded reference model module."""
    ns = types.SimpleNamespace
    if "diffusers" not in sys.modules:
        for name in ("diffusers", "diffusers.configuration_utils", "diffusers.models", "diffusers.models.modeling_utils"):
            sys.modules[name] = types.ModuleType(name)

        class ConfigMixin:  # plumbing only
            pass
        sys.modules["diffusers.configuration_utils"].ConfigMixin = ConfigMixin
        sys.modules["diffusers.configuration_utils"].register_to_config = lambda f: f
        sys.modules["diffusers.models.modeling_utils"].ModelMixin = torch.nn.Module
    for pk in (tree, f"{tree}.modules"):
        if pk not in sys.modules:
            m = types.ModuleType(pk)
            m.__path__ = [str(REF / pk.replace(".", "/"))]
            sys.modules[pk] = m

    def load(modname, path):
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    att = load(f"{tree}.modules.attention", REF / tree / "modules" / "attention.py")
    att.flash_attention = _sdpa_flash_attention
    return load(f"{tree}.modules.model", REF / tree / "modules" / "model.py")


def build_reference_model(mod, cfg: dict, sd: dict):
    kw = dict(model_type="ti2v" if cfg["variant"] == "5b" else "i2v", text_len=cfg["text_len"], in_dim=cfg["in_dim"],
              dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"],
              out_dim=cfg["out_dim"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], eps=1e-6)
    model = mod.WanModel(**kw)
    if cfg["variant"] == "14b":
        # what wan/image2video.py:155-159 attaches (shapes only matter; the values come from the state dict)
        cin, C = cfg["in_dim"], cfg["dim"]
        for name, k in (("patch_embedding_2x", 4), ("patch_embedding_4x", 8), ("patch_embedding_8x", 16),
                        ("patch_embedding_16x", 32)):
            setattr(model, name, torch.nn.Conv3d(cin, C, (1, k, k), (1, k, k)))
        model.patch_embedding_2x_f = torch.nn.Conv3d(cin, cin, (1, 4, 4), (1, 4, 4))
        # MLPProj is hard-wired to 1280 inputs (wan/modules/model.py:669); rebuild it at the tiny clip width
        if cfg["clip_dim"] != 1280:
            model.img_emb = mod.MLPProj(cfg["clip_dim"], C)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.eval()


def cases_5b():
    # (name, frames, H, W, latent_frame_zero, flag, seq_len_pad, t)
    return [
        ("5b_grid", 3, 8, 8, None, False, 0, [500.0]),
        ("5b_grid_padded", 2, 8, 12, None, False, 16, [250.0]),
        ("5b_pack_h3", 3 + 2, 6, 10, 2, True, 0, [[0.0, 900.0]]),
        ("5b_pack_h1", 1 + 2, 6, 10, 2, True, 0, [[0.0, 700.0]]),
        ("5b_pack_h10", 10 + 2, 6, 10, 2, True, 0, [[0.0, 400.0]]),
        ("5b_pack_h30", 30 + 2, 6, 10, 2, True, 0, [[0.0, 999.0]]),
        ("5b_pack_h100", 100 + 1, 6, 10, 1, True, 0, [[0.0, 100.0]]),
        ("5b_pack_h400", 400 + 1, 6, 10, 1, True, 0, [[0.0, 650.0]]),
    ]


def cases_14b():
    # (name, frames, H, W, latent_frame_zero, rand_num_img, t)
    return [
        ("14b_grid", 3, 8, 8, 9, 0.2, [300.0]),
        ("14b_pack_h4", 4 + 9, 6, 10, 9, 0.6, [800.0]),
        ("14b_pack_lfz8", 5 + 8, 6, 10, 8, 0.6, [600.0]),
        ("14b_pack_h12", 12 + 9, 6, 10, 9, 0.6, [450.0]),
    ]


def cases_h8():
    """8-head models (dim 1024): Ulysses parity at world 2 / 4 / 8 (tools/sp_parity.py). Token counts are chosen NOT to
    divide by 8 so the shard padding path runs; 14b_grid_padded has seq_len > F*H*W (k_lens masking,
    wan/modules/model.py:311-314)."""
    five = [("5b_grid", 3, 10, 14, None, False, 0, [500.0]),          # L = 105
            ("5b_grid_padded", 2, 10, 14, None, False, 9, [250.0]),   # L = 70 + 9 padded rows (keys on the 5B tree)
            ("5b_pack_h10", 10 + 2, 10, 14, 2, True, 0, [[0.0, 400.0]]),
            ("5b_pack_h30", 30 + 2, 6, 10, 2, True, 0, [[0.0, 999.0]])]
    fourteen = [("14b_grid", 3, 10, 14, 9, 0.2, 0, [300.0]),
                ("14b_grid_padded", 3, 10, 14, 9, 0.2, 7, [300.0]),   # 105 tokens, seq_len 112: 7 masked padding rows
                ("14b_pack_lfz8", 5 + 8, 6, 10, 8, 0.6, 0, [600.0])]
    return five, fourteen


@torch.no_grad()
def main_h8(mod23, mod21, out_dir, seed_w):
    five, fourteen = cases_h8()
    cfg = synth.CFG_5B_H8
    sd = synth.make_state_dict(cfg, seed_w + 10)
    model = build_reference_model(mod23, cfg, sd)
    gold = {"cfg": cfg, "seed_w": seed_w + 10, "cases": {}}
    for i, (name, frames, H, W, lfz, flag, pad, t) in enumerate(five):
        inp = synth.make_inputs(cfg, 300 + i, frames, H, W, ctx_len=20)
        L_grid = frames * (H // 2) * (W // 2)
        kwargs = dict(seq_len=L_grid + pad, flag=flag)
        if lfz is not None:
            kwargs["latent_frame_zero"] = lfz
        out = model([inp["x"]], torch.tensor(t), [inp["context"]], **kwargs)[0]
        gold["cases"][name] = dict(seed=300 + i, frames=frames, H=H, W=W, lfz=lfz, flag=flag, seq_len=L_grid + pad, t=t,
                                   ctx_len=20, out=out.clone())
        print("h8", name, tuple(out.shape), float(out.abs().mean()))
    torch.save(gold, out_dir / "wan23_h8.pt")
    cfg = synth.CFG_14B_H8
    sd = synth.make_state_dict(cfg, seed_w + 11)
    model = build_reference_model(mod21, cfg, sd)
    gold = {"cfg": cfg, "seed_w": seed_w + 11, "cases": {}}
    for i, (name, frames, H, W, lfz, rni, pad, t) in enumerate(fourteen):
        inp = synth.make_inputs(cfg, 400 + i, frames, H, W, ctx_len=20)
        L_grid = frames * (H // 2) * (W // 2)
        out, _ = model([inp["x"]], torch.tensor(t), [inp["context"]], seq_len=L_grid + pad, clip_fea=inp["clip_fea"],
                       y=[inp["y"]], rand_num_img=rni, latent_frame_zero=lfz)
        gold["cases"][name] = dict(seed=400 + i, frames=frames, H=H, W=W, lfz=lfz, rand_num_img=rni, seq_len=L_grid + pad,
                                   t=t, ctx_len=20, out=out.clone())
        print("h8", name, tuple(out.shape), float(out.abs().mean()))
    torch.save(gold, out_dir / "wan21_h8.pt")


@torch.no_grad()
def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    seed_w = 1234

    # ---------------- 5B tree ----------------
    mod23 = load_reference("wan23")
    cfg = synth.CFG_5B_TINY
    sd = synth.make_state_dict(cfg, seed_w)
    model = build_reference_model(mod23, cfg, sd)
    gold = {"cfg": cfg, "seed_w": seed_w, "weight_abs_sum": float(sum(v.abs().sum() for v in sd.values())), "cases": {}}
    for i, (name, frames, H, W, lfz, flag, pad, t) in enumerate(cases_5b()):
        inp = synth.make_inputs(cfg, 100 + i, frames, H, W, ctx_len=20)
        L_grid = frames * (H // 2) * (W // 2)
        kwargs = dict(seq_len=L_grid + pad, flag=flag)
        if lfz is not None:
            kwargs["latent_frame_zero"] = lfz
        out = model([inp["x"]], torch.tensor(t), [inp["context"]], **kwargs)[0]
        gold["cases"][name] = dict(seed=100 + i, frames=frames, H=H, W=W, lfz=lfz, flag=flag, seq_len=L_grid + pad, t=t,
                                   ctx_len=20, out=out.clone())
        print(name, tuple(out.shape), float(out.abs().mean()))
    # single-block fixture (BASELINE.json configs[0]: one WanAttentionBlock, 128 tokens, grid 2x8x8)
    g = torch.Generator().manual_seed(7)
    Lb, C = 128, cfg["dim"]
    xb = torch.randn(1, Lb, C, generator=g)
    eb = 0.5 * torch.randn(1, Lb, 6, C, generator=g)
    cb = torch.randn(1, cfg["text_len"], C, generator=g)
    d = cfg["dim"] // cfg["num_heads"]  # the [1024, d/2] grid table of model.py:475-480 (model.freqs was overwritten
    freqs = torch.cat([mod23.rope_params(1024, d - 4 * (d // 6)), mod23.rope_params(1024, 2 * (d // 6)),  # by forward)
                       mod23.rope_params(1024, 2 * (d // 6))], dim=1)
    yb = model.blocks[0](xb, eb, torch.tensor([Lb]), torch.tensor([[2, 8, 8]]), freqs, cb, None, flag=False)
    gold["block"] = dict(seed=7, L=Lb, grid=(2, 8, 8), out=yb.clone())
    torch.save(gold, out_dir / "wan23_tiny.pt")

    # ---------------- 14B tree ----------------
    mod21 = load_reference("wan")
    cfg = synth.CFG_14B_TINY
    sd = synth.make_state_dict(cfg, seed_w + 1)
    model = build_reference_model(mod21, cfg, sd)
    gold = {"cfg": cfg, "seed_w": seed_w + 1, "weight_abs_sum": float(sum(v.abs().sum() for v in sd.values())), "cases": {}}
    for i, (name, frames, H, W, lfz, rni, t) in enumerate(cases_14b()):
        inp = synth.make_inputs(cfg, 200 + i, frames, H, W, ctx_len=20)
        L_grid = frames * (H // 2) * (W // 2)
        out, _ = model([inp["x"]], torch.tensor(t), [inp["context"]], seq_len=L_grid, clip_fea=inp["clip_fea"],
                       y=[inp["y"]], rand_num_img=rni, latent_frame_zero=lfz)
        gold["cases"][name] = dict(seed=200 + i, frames=frames, H=H, W=W, lfz=lfz, rand_num_img=rni, seq_len=L_grid, t=t,
                                   ctx_len=20, out=out.clone())
        print(name, tuple(out.shape), float(out.abs().mean()))
    torch.save(gold, out_dir / "wan21_tiny.pt")
    main_h8(mod23, mod21, out_dir, seed_w)
    for f in sorted(out_dir.glob("*.pt")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
