"""CPU oracle for the Wan2.1 VAE *encode* path (`WanVAE.encode`, the conditioning encode of the 14B I2V sampler,
`wan/image2video.py:348-367`) — TEST INFRASTRUCTURE ONLY.

Whole-sequence restatement of /root/reference/wan/modules/vae.py `WanVAE_.encode` (:515-542: frame 0, then 4 frames per
`Encoder3d.forward` (:265-366) call with a feature cache), same unrolling as oracle/wan22vae_enc.py: causal zero-padded
convs over the whole sequence; `Resample(downsample3d)` (:84-90, :125-139) lets frame 0 pass and runs a stride-2 unpadded
temporal conv over the rest. Differences from 2.2: flat `encoder.downsamples` Sequential (:293-306), no AvgDown3D
shortcut, no patchify (RGB straight into `encoder.conv1`).
tests/golden/wan21vae_enc_tiny.pt (tools/make_golden_vae21_enc.py) comes from the reference's own chunked encode.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .wan22vae import causal_conv3d, rms_norm
from .wan22vae_enc import Wan22VaeEncodeOracle

Tensor = torch.Tensor


def layer_plan(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True)) -> List[Tuple]:
    """The flat `encoder.downsamples` Sequential (:293-306) as (index, kind, in_dim, out_dim)."""
    dims = [dim * u for u in [1] + list(dim_mult)]
    plan, n = [], 0
    for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            plan.append((n, "res", ci, co))
            n, ci = n + 1, co
        if i != len(dim_mult) - 1:
            plan.append((n, "downsample3d" if temperal_downsample[i] else "downsample2d", co, co))
            n += 1
    return plan


class Wan21VaeEncodeOracle(Wan22VaeEncodeOracle):
    def __init__(self, sd: Dict[str, Tensor], dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_downsample=(False, True, True), mean=None, std=None, **_):
        self.sd, self.z_dim = sd, z_dim
        self.plan = layer_plan(dim, dim_mult, num_res_blocks, temperal_downsample)
        self.mean = torch.zeros(z_dim) if mean is None else mean
        self.std = torch.ones(z_dim) if std is None else std

    @torch.no_grad()
    def encode(self, video: Tensor) -> Tensor:
        """WanVAE.encode (:645-653) for one video [3, T, H, W] (T = 1 + 4k) -> mu [z_dim, 1 + k, H/8, W/8]."""
        x = self._conv("encoder.conv1", video.unsqueeze(0).float())
        for n, kind, _, _ in self.plan:
            p = f"encoder.downsamples.{n}"
            x = self.res_block(p, x) if kind == "res" else self.resample_down(p, x, kind == "downsample3d")
        x = self.res_block("encoder.middle.0", x)
        x = self.attn_block("encoder.middle.1", x)
        x = self.res_block("encoder.middle.2", x)
        x = self._conv("encoder.head.2", F.silu(rms_norm(x, self.sd["encoder.head.0.gamma"])))
        mu = self._conv("conv1", x)[:, :self.z_dim]
        mu = (mu - self.mean.view(1, -1, 1, 1, 1)) / self.std.view(1, -1, 1, 1, 1)
        return mu.float().squeeze(0)


def param_shapes(dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True)):
    d0 = dim
    s: Dict[str, tuple] = {"conv1.weight": (2 * z_dim, 2 * z_dim, 1, 1, 1), "conv1.bias": (2 * z_dim,),
                           "encoder.conv1.weight": (d0, 3, 3, 3, 3), "encoder.conv1.bias": (d0,)}

    def res(p, ci, co):
        s[p + ".residual.0.gamma"] = (ci, 1, 1, 1)
        s[p + ".residual.2.weight"], s[p + ".residual.2.bias"] = (co, ci, 3, 3, 3), (co,)
        s[p + ".residual.3.gamma"] = (co, 1, 1, 1)
        s[p + ".residual.6.weight"], s[p + ".residual.6.bias"] = (co, co, 3, 3, 3), (co,)
        if ci != co:
            s[p + ".shortcut.weight"], s[p + ".shortcut.bias"] = (co, ci, 1, 1, 1), (co,)

    last = d0
    for n, kind, ci, co in layer_plan(dim, dim_mult, num_res_blocks, temperal_downsample):
        p = f"encoder.downsamples.{n}"
        if kind == "res":
            res(p, ci, co)
            last = co
        else:
            s[p + ".resample.1.weight"], s[p + ".resample.1.bias"] = (co, ci, 3, 3), (co,)
            if kind == "downsample3d":
                s[p + ".time_conv.weight"], s[p + ".time_conv.bias"] = (co, ci, 3, 1, 1), (co,)
    res("encoder.middle.0", last, last)
    s["encoder.middle.1.norm.gamma"] = (last, 1, 1)
    s["encoder.middle.1.to_qkv.weight"], s["encoder.middle.1.to_qkv.bias"] = (3 * last, last, 1, 1), (3 * last,)
    s["encoder.middle.1.proj.weight"], s["encoder.middle.1.proj.bias"] = (last, last, 1, 1), (last,)
    res("encoder.middle.2", last, last)
    s["encoder.head.0.gamma"] = (last, 1, 1, 1)
    s["encoder.head.2.weight"], s["encoder.head.2.bias"] = (2 * z_dim, last, 3, 3, 3), (2 * z_dim,)
    return s


def make_state_dict(seed: int, **cfg) -> Dict[str, Tensor]:
    sd = {}
    for idx, (name, shape) in enumerate(param_shapes(**cfg).items()):
        g = torch.Generator().manual_seed(seed * 6007 + idx)
        t = torch.randn(shape, generator=g)
        if name.endswith(".bias"):
            t = 0.05 * t
        elif name.endswith(".gamma"):
            t = 1.0 + 0.1 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (1.0 / fan_in ** 0.5)
        sd[name] = t.to(torch.bfloat16).float()
    return sd
