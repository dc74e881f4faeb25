"""Generate tests/golden/hyvae_tiny.pt by running the REFERENCE's hyvideo/vae code (imported from /root/reference) on
CPU at reduced width. Authoring container only. `diffusers` is absent: everything imported from it by
hyvideo/vae/*.py is plumbing and stubbed, except `diffusers.models.attention_processor.Attention`, whose arithmetic
is restated below for the constructor arguments used at unet_causal_3d_blocks.py:580-592."""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
from oracle import hyvae  # noqa: E402


class Attention(nn.Module):
    """diffusers==0.32.0 Attention for (heads=1, norm_num_groups=32, residual_connection=True, bias=True,
    upcast_softmax=True, _from_deprecated_attn_block=True, rescale_output_factor=1)."""

    def __init__(self, query_dim, heads=1, dim_head=64, rescale_output_factor=1.0, eps=1e-5, norm_num_groups=None,
                 spatial_norm_dim=None, residual_connection=False, bias=False, upcast_softmax=False,
                 _from_deprecated_attn_block=False, **kw):
        super().__init__()
        inner = heads * dim_head
        assert heads == 1 and spatial_norm_dim is None
        self.scale, self.rescale, self.residual = dim_head ** -0.5, rescale_output_factor, residual_connection
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True)
        self.to_q, self.to_k, self.to_v = (nn.Linear(query_dim, inner, bias=bias) for _ in range(3))
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states, temb=None, attention_mask=None):
        residual = hidden_states
        h = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        scores = torch.baddbmm(attention_mask.to(q.dtype), q, k.transpose(1, 2), beta=1, alpha=self.scale)
        probs = scores.float().softmax(dim=-1).to(q.dtype)
        o = self.to_out[1](self.to_out[0](torch.bmm(probs, v)))
        if self.residual:
            o = o + residual
        return o / self.rescale


def load_reference_vae():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class ConfigMixin:
        pass

    def register_to_config(init):
        def wrapped(self, *a, **kw):
            import inspect
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            self.config = types.SimpleNamespace(**cfg)
            init(self, *a, **kw)
        return wrapped

    class BaseOutput(dict):
        def __post_init__(self):
            pass

    class _Logger:
        def __getattr__(self, n):
            return lambda *a, **k: None

    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    mod("diffusers.models")
    mod("diffusers.models.modeling_utils", ModelMixin=nn.Module)
    mod("diffusers.models.modeling_outputs", AutoencoderKLOutput=object)
    mod("diffusers.models.activations", get_activation=lambda n: {"silu": nn.SiLU(), "swish": nn.SiLU()}[n])
    mod("diffusers.models.normalization", AdaGroupNorm=object, RMSNorm=object)
    names = ["ADDED_KV_ATTENTION_PROCESSORS", "CROSS_ATTENTION_PROCESSORS", "AttentionProcessor", "AttnAddedKVProcessor",
             "AttnProcessor"]
    mod("diffusers.models.attention_processor", Attention=Attention, SpatialNorm=object, **{n: object for n in names})
    mod("diffusers.utils", BaseOutput=BaseOutput, is_torch_version=lambda *a: True,
        logging=types.SimpleNamespace(get_logger=lambda n: _Logger()))
    mod("diffusers.utils.torch_utils", randn_tensor=None)
    mod("diffusers.utils.accelerate_utils", apply_forward_hook=lambda f: f)
    mod("diffusers.loaders", FromOriginalVAEMixin=object)
    for pk in ("hyvideo", "hyvideo.vae"):
        m = types.ModuleType(pk)
        m.__path__ = [str(REF / pk.replace(".", "/"))]
        sys.modules[pk] = m

    def load(modname, path):
        spec = importlib.util.spec_from_file_location(modname, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m

    load("hyvideo.vae.unet_causal_3d_blocks", REF / "hyvideo/vae/unet_causal_3d_blocks.py")
    load("hyvideo.vae.vae", REF / "hyvideo/vae/vae.py")
    return load("hyvideo.vae.autoencoder_kl_causal_3d", REF / "hyvideo/vae/autoencoder_kl_causal_3d.py")


TINY = dict(block_out_channels=(64, 64, 128, 128), layers_per_block=2, latent_channels=16, out_channels=3)


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    ak = load_reference_vae()
    seed = 4321
    sd = hyvae.make_state_dict(seed, **TINY)
    gold = {"cfg": TINY, "seed_w": seed, "weight_abs_sum": float(sum(v.abs().sum() for v in sd.values())), "cases": {}}
    cases = [
        # name, z shape (T,H,W), sample_size, sample_tsize, tiling
        ("untiled", (3, 8, 8), 256, 64, False),
        ("untiled_t1", (1, 6, 10), 256, 64, False),
        ("spatial_tiled", (2, 12, 14), 64, 64, True),
        ("temporal_spatial_tiled", (6, 10, 12), 64, 16, True),
    ]
    for i, (name, (T, H, W), ss, st, tiling) in enumerate(cases):
        vae = ak.AutoencoderKLCausal3D(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlockCausal3D",) * 4,
                                       up_block_types=("UpDecoderBlockCausal3D",) * 4,
                                       block_out_channels=TINY["block_out_channels"], layers_per_block=2, act_fn="silu",
                                       latent_channels=16, norm_num_groups=32, sample_size=ss, sample_tsize=st,
                                       time_compression_ratio=4, spatial_compression_ratio=8, mid_block_add_attention=True)
        missing, unexpected = vae.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith(("encoder.", "quant_conv.")) for k in missing), (missing[:5], unexpected)
        vae.eval()
        if tiling:
            vae.enable_tiling()
        g = torch.Generator().manual_seed(300 + i)
        z = torch.randn(1, 16, T, H, W, generator=g)
        out = vae.decode(z, return_dict=False)[0]
        # fixtures stay small: a stride-3 sample of the pixels plus full row / column sums (every pixel contributes)
        gold["cases"][name] = dict(seed=300 + i, T=T, H=H, W=W, sample_size=ss, sample_tsize=st, tiling=tiling,
                                   shape=tuple(out.shape), sample=out[..., ::3, ::3].clone(), rowsum=out.sum(-1),
                                   colsum=out.sum(-2))
        print(name, tuple(out.shape), float(out.abs().mean()))
    path = ROOT / "tests" / "golden" / "hyvae_tiny.pt"
    torch.save(gold, path)
    print(path.name, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
