"""CPU oracle for the Wan2.1 VAE *decode* path (`WanVAE.decode`, what the 14B sampler calls, wan/image2video.py:197) —
TEST INFRASTRUCTURE ONLY.

Restates /root/reference/wan/modules/vae.py (identical to wan23/modules/vae2_1.py) in WHOLE-SEQUENCE form, exactly as
oracle/wan22vae.py does for the 2.2 VAE: `WanVAE_.decode` (:544-568) pushes one latent frame per call through
`Decoder3d.forward` (:421-472) with a feature cache in every CausalConv3d (:17-36, `ResidualBlock.forward` :207-224,
`Resample.forward` :100-139); unrolled, every conv is a causal conv over the whole sequence with zero padding, frame 0
bypasses `time_conv` ("Rep", :104-106) and `time_conv` runs over frames 1.. with zero history (:108-131).
Differences from the 2.2 decoder: the flat `decoder.upsamples` Sequential (:395-414), `Resample`'s Conv2d halves the
channel count (:76-83), blocks 1..3 therefore start from dims[i] // 2 (:398-399), no DupUp3D shortcut, and the head conv
emits RGB directly (:417-419) — no unpatchify.
tests/golden/wan21vae_tiny.pt (tools/make_golden_vae21.py) comes from the reference's own chunked code.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .wan22vae import causal_conv3d, rms_norm

Tensor = torch.Tensor


def layer_plan(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_upsample=(True, True, False)) -> List[Tuple]:
    """The flat `decoder.upsamples` Sequential (:395-414) as (index, kind, in_dim, out_dim) with kind in
    res / upsample2d / upsample3d."""
    dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    plan, n = [], 0
    for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            ci = ci // 2
        for _ in range(num_res_blocks + 1):
            plan.append((n, "res", ci, co))
            n += 1
            ci = co
        if i != len(dim_mult) - 1:
            plan.append((n, "upsample3d" if temperal_upsample[i] else "upsample2d", co, co // 2))
            n += 1
    return plan


class Wan21VaeOracle:
    def __init__(self, sd: Dict[str, Tensor], dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_upsample=(True, True, False), mean=None, std=None, **_):
        self.sd, self.z_dim = sd, z_dim
        self.plan = layer_plan(dim, dim_mult, num_res_blocks, temperal_upsample)
        self.mean = torch.zeros(z_dim) if mean is None else mean
        self.std = torch.ones(z_dim) if std is None else std

    def _conv(self, p: str, x: Tensor) -> Tensor:
        return causal_conv3d(x, self.sd[p + ".weight"], self.sd[p + ".bias"])

    def res_block(self, p: str, x: Tensor) -> Tensor:
        """ResidualBlock (:186-224)."""
        h = self._conv(p + ".shortcut", x) if (p + ".shortcut.weight") in self.sd else x
        y = self._conv(p + ".residual.2", F.silu(rms_norm(x, self.sd[p + ".residual.0.gamma"])))
        y = self._conv(p + ".residual.6", F.silu(rms_norm(y, self.sd[p + ".residual.3.gamma"])))
        return y + h

    def attn_block(self, p: str, x: Tensor) -> Tensor:
        """AttentionBlock (:227-262)."""
        b, c, t, h, w = x.shape
        y = rms_norm(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), self.sd[p + ".norm.gamma"])
        qkv = F.conv2d(y, self.sd[p + ".to_qkv.weight"], self.sd[p + ".to_qkv.bias"])
        q, k, v = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
        o = F.scaled_dot_product_attention(q, k, v).squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
        o = F.conv2d(o, self.sd[p + ".proj.weight"], self.sd[p + ".proj.bias"])
        return o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x

    def resample(self, p: str, x: Tensor, temporal: bool) -> Tensor:
        """Resample upsample2d / upsample3d (:61-139) over the whole sequence; the Conv2d halves the channels."""
        b, c, t, h, w = x.shape
        if temporal and t > 1:
            y = self._conv(p + ".time_conv", x[:, :, 1:]).reshape(b, 2, c, t - 1, h, w)
            y = torch.stack((y[:, 0], y[:, 1]), 3).reshape(b, c, 2 * (t - 1), h, w)
            x = torch.cat([x[:, :, :1], y], dim=2)
        t = x.shape[2]
        y = F.interpolate(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w).float(), scale_factor=(2.0, 2.0), mode="nearest-exact")
        y = F.conv2d(y, self.sd[p + ".resample.1.weight"], self.sd[p + ".resample.1.bias"], padding=1)
        return y.reshape(b, t, y.shape[1], 2 * h, 2 * w).permute(0, 2, 1, 3, 4)

    @torch.no_grad()
    def decode(self, z: Tensor) -> Tensor:
        """WanVAE.decode (:655-663) for one latent z [z_dim, T, H, W] -> [3, 4(T-1)+1, 8H, 8W] in [-1, 1]."""
        z = z.unsqueeze(0).float()
        z = z * self.std.view(1, -1, 1, 1, 1) + self.mean.view(1, -1, 1, 1, 1)       # z / (1/std) + mean (:547-551)
        x = self._conv("decoder.conv1", self._conv("conv2", z))
        x = self.res_block("decoder.middle.0", x)
        x = self.attn_block("decoder.middle.1", x)
        x = self.res_block("decoder.middle.2", x)
        for n, kind, _, _ in self.plan:
            p = f"decoder.upsamples.{n}"
            x = self.res_block(p, x) if kind == "res" else self.resample(p, x, kind == "upsample3d")
        x = self._conv("decoder.head.2", F.silu(rms_norm(x, self.sd["decoder.head.0.gamma"])))
        return x.float().clamp_(-1, 1).squeeze(0)


def param_shapes(dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_upsample=(True, True, False)):
    dims0 = dim * dim_mult[-1]
    s: Dict[str, tuple] = {"conv2.weight": (z_dim, z_dim, 1, 1, 1), "conv2.bias": (z_dim,),
                           "decoder.conv1.weight": (dims0, z_dim, 3, 3, 3), "decoder.conv1.bias": (dims0,)}

    def res(p, ci, co):
        s[p + ".residual.0.gamma"] = (ci, 1, 1, 1)
        s[p + ".residual.2.weight"], s[p + ".residual.2.bias"] = (co, ci, 3, 3, 3), (co,)
        s[p + ".residual.3.gamma"] = (co, 1, 1, 1)
        s[p + ".residual.6.weight"], s[p + ".residual.6.bias"] = (co, co, 3, 3, 3), (co,)
        if ci != co:
            s[p + ".shortcut.weight"], s[p + ".shortcut.bias"] = (co, ci, 1, 1, 1), (co,)

    res("decoder.middle.0", dims0, dims0)
    s["decoder.middle.1.norm.gamma"] = (dims0, 1, 1)
    s["decoder.middle.1.to_qkv.weight"], s["decoder.middle.1.to_qkv.bias"] = (3 * dims0, dims0, 1, 1), (3 * dims0,)
    s["decoder.middle.1.proj.weight"], s["decoder.middle.1.proj.bias"] = (dims0, dims0, 1, 1), (dims0,)
    res("decoder.middle.2", dims0, dims0)
    out_dim = dims0
    for n, kind, ci, co in layer_plan(dim, dim_mult, num_res_blocks, temperal_upsample):
        p = f"decoder.upsamples.{n}"
        if kind == "res":
            res(p, ci, co)
            out_dim = co
        else:
            s[p + ".resample.1.weight"], s[p + ".resample.1.bias"] = (co, ci, 3, 3), (co,)
            if kind == "upsample3d":
                s[p + ".time_conv.weight"], s[p + ".time_conv.bias"] = (2 * ci, ci, 3, 1, 1), (2 * ci,)
    s["decoder.head.0.gamma"] = (out_dim, 1, 1, 1)
    s["decoder.head.2.weight"], s["decoder.head.2.bias"] = (3, out_dim, 3, 3, 3), (3,)
    return s


def make_state_dict(seed: int, **cfg) -> Dict[str, Tensor]:
    sd = {}
    for idx, (name, shape) in enumerate(param_shapes(**cfg).items()):
        g = torch.Generator().manual_seed(seed * 7919 + idx)
        t = torch.randn(shape, generator=g)
        if name.endswith(".bias"):
            t = 0.05 * t
        elif name.endswith(".gamma"):
            t = 1.0 + 0.1 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (0.8 / fan_in ** 0.5)
        sd[name] = t.to(torch.bfloat16).float()
    return sd
