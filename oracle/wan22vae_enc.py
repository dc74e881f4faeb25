"""CPU oracle for the Wan2.2 VAE *encode* path (`Wan2_2_VAE.encode`, SURVEY.md §8(f) rank 3: the AR history / first-frame
conditioning encode, `sample_5b.py:892-893`) — TEST INFRASTRUCTURE ONLY.

Restates /root/reference/wan23/modules/vae2_2.py in WHOLE-SEQUENCE form, like oracle/wan22vae.py does for decode. The
reference encodes frame 0 alone and then 4 frames per `Encoder3d.forward` call with a per-conv feature cache
(`WanVAE_.encode` :796-829). Unrolled:
  * every CausalConv3d is a causal conv over the whole sequence with zero padding in front (:22-44, :216-239, :566-620);
  * `Resample(downsample3d)` (:105-110, :158-170): frame 0 bypasses `time_conv`; the remaining frames go through a
    stride-2, kernel-3, unpadded temporal conv over [f0, f1, f2, ...] — outputs conv(f_2k, f_2k+1, f_2k+2);
  * `AvgDown3D` (:320-373) pads `factor_t - T % factor_t` zero frames in FRONT per call; because every chunked call sees
    1 frame first and an even count afterwards this equals padding the whole (odd-length) sequence once.
tests/golden/wan22vae_enc_tiny.pt (tools/make_golden_vae22_enc.py) is produced by the reference's own chunked encode.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .wan22vae import causal_conv3d, rms_norm

Tensor = torch.Tensor


def patchify2(x: Tensor) -> Tensor:
    """patchify(patch_size=2) 'b c f (h q) (w r) -> b (c r q) f h w' (:284-300)."""
    b, c, f, H, W = x.shape
    x = x.view(b, c, f, H // 2, 2, W // 2, 2)                    # b c f h q w r
    return x.permute(0, 1, 6, 4, 2, 3, 5).reshape(b, c * 4, f, H // 2, W // 2)


def avg_down(x: Tensor, out_c: int, ft: int, fs: int) -> Tensor:
    """AvgDown3D (:320-373) on the whole sequence."""
    pad_t = (ft - x.shape[2] % ft) % ft
    x = F.pad(x, (0, 0, 0, 0, pad_t, 0))
    B, C, T, H, W = x.shape
    x = x.view(B, C, T // ft, ft, H // fs, fs, W // fs, fs).permute(0, 1, 3, 5, 7, 2, 4, 6).contiguous()
    x = x.view(B, out_c, C * ft * fs * fs // out_c, T // ft, H // fs, W // fs)
    return x.mean(dim=2)


class Wan22VaeEncodeOracle:
    def __init__(self, sd: Dict[str, Tensor], dim=160, z_dim=48, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_downsample=(False, True, True), mean=None, std=None, **_):
        self.sd, self.z_dim, self.nrb = sd, z_dim, num_res_blocks
        self.dims = [dim * u for u in [1] + list(dim_mult)]                      # :527
        self.t_down = list(temperal_downsample)
        self.n = len(dim_mult)
        self.mean = torch.zeros(z_dim) if mean is None else mean
        self.std = torch.ones(z_dim) if std is None else std

    def _conv(self, p: str, x: Tensor) -> Tensor:
        return causal_conv3d(x, self.sd[p + ".weight"], self.sd[p + ".bias"])

    def res_block(self, p: str, x: Tensor) -> Tensor:
        h = self._conv(p + ".shortcut", x) if (p + ".shortcut.weight") in self.sd else x
        y = self._conv(p + ".residual.2", F.silu(rms_norm(x, self.sd[p + ".residual.0.gamma"])))
        y = self._conv(p + ".residual.6", F.silu(rms_norm(y, self.sd[p + ".residual.3.gamma"])))
        return y + h

    def attn_block(self, p: str, x: Tensor) -> Tensor:
        b, c, t, h, w = x.shape
        y = rms_norm(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), self.sd[p + ".norm.gamma"])
        qkv = F.conv2d(y, self.sd[p + ".to_qkv.weight"], self.sd[p + ".to_qkv.bias"])
        q, k, v = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
        o = F.scaled_dot_product_attention(q, k, v).squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
        o = F.conv2d(o, self.sd[p + ".proj.weight"], self.sd[p + ".proj.bias"])
        return o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x

    def resample_down(self, p: str, x: Tensor, temporal: bool) -> Tensor:
        """Resample downsample2d / downsample3d (:101-110, :152-170): ZeroPad2d(0,1,0,1) + Conv2d 3x3 stride 2 per frame,
        then (3d) frame 0 passes and the rest is a stride-2 temporal conv over the whole sequence."""
        b, c, t, h, w = x.shape
        y = F.pad(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), (0, 1, 0, 1))
        y = F.conv2d(y, self.sd[p + ".resample.1.weight"], self.sd[p + ".resample.1.bias"], stride=2)
        y = y.reshape(b, t, c, y.shape[-2], y.shape[-1]).permute(0, 2, 1, 3, 4)
        if temporal and t > 1:
            z = F.conv3d(y, self.sd[p + ".time_conv.weight"], self.sd[p + ".time_conv.bias"], stride=(2, 1, 1))
            y = torch.cat([y[:, :, :1], z], dim=2)
        return y

    def down_block(self, i: int, x: Tensor) -> Tensor:
        """Down_ResidualBlock (:420-459)."""
        p = f"encoder.downsamples.{i}.downsamples"
        down = i != self.n - 1
        t_down = self.t_down[i] if i < len(self.t_down) else False
        short = avg_down(x, self.dims[i + 1], 2 if t_down else 1, 2 if down else 1)
        for j in range(self.nrb):
            x = self.res_block(f"{p}.{j}", x)
        if down:
            x = self.resample_down(f"{p}.{self.nrb}", x, t_down)
        return x + short

    @torch.no_grad()
    def encode(self, video: Tensor) -> Tensor:
        """Wan2_2_VAE.encode (:1042-1057) for one video [3, T, H, W] (T = 1 + 4k) -> mu [z_dim, 1 + k, H/16, W/16]."""
        x = patchify2(video.unsqueeze(0).float())
        x = self._conv("encoder.conv1", x)
        for i in range(self.n):
            x = self.down_block(i, x)
        x = self.res_block("encoder.middle.0", x)
        x = self.attn_block("encoder.middle.1", x)
        x = self.res_block("encoder.middle.2", x)
        x = self._conv("encoder.head.2", F.silu(rms_norm(x, self.sd["encoder.head.0.gamma"])))
        mu = self._conv("conv1", x)[:, :self.z_dim]                                # .chunk(2, dim=1)[0] (:822)
        mu = (mu - self.mean.view(1, -1, 1, 1, 1)) / self.std.view(1, -1, 1, 1, 1)    # (mu - mean) * (1/std) (:823-828)
        return mu.float().squeeze(0)


def param_shapes(dim=160, z_dim=48, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True)):
    dims = [dim * u for u in [1] + list(dim_mult)]
    s: Dict[str, tuple] = {"conv1.weight": (2 * z_dim, 2 * z_dim, 1, 1, 1), "conv1.bias": (2 * z_dim,),
                           "encoder.conv1.weight": (dims[0], 12, 3, 3, 3), "encoder.conv1.bias": (dims[0],)}

    def res(p, ci, co):
        s[p + ".residual.0.gamma"] = (ci, 1, 1, 1)
        s[p + ".residual.2.weight"], s[p + ".residual.2.bias"] = (co, ci, 3, 3, 3), (co,)
        s[p + ".residual.3.gamma"] = (co, 1, 1, 1)
        s[p + ".residual.6.weight"], s[p + ".residual.6.bias"] = (co, co, 3, 3, 3), (co,)
        if ci != co:
            s[p + ".shortcut.weight"], s[p + ".shortcut.bias"] = (co, ci, 1, 1, 1), (co,)

    n = len(dim_mult)
    for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        p, c = f"encoder.downsamples.{i}.downsamples", ci
        for j in range(num_res_blocks):
            res(f"{p}.{j}", c, co)
            c = co
        if i != n - 1:
            q = f"{p}.{num_res_blocks}"
            s[q + ".resample.1.weight"], s[q + ".resample.1.bias"] = (co, co, 3, 3), (co,)
            if i < len(temperal_downsample) and temperal_downsample[i]:
                s[q + ".time_conv.weight"], s[q + ".time_conv.bias"] = (co, co, 3, 1, 1), (co,)
    d = dims[-1]
    res("encoder.middle.0", d, d)
    s["encoder.middle.1.norm.gamma"] = (d, 1, 1)
    s["encoder.middle.1.to_qkv.weight"], s["encoder.middle.1.to_qkv.bias"] = (3 * d, d, 1, 1), (3 * d,)
    s["encoder.middle.1.proj.weight"], s["encoder.middle.1.proj.bias"] = (d, d, 1, 1), (d,)
    res("encoder.middle.2", d, d)
    s["encoder.head.0.gamma"] = (d, 1, 1, 1)
    s["encoder.head.2.weight"], s["encoder.head.2.bias"] = (2 * z_dim, d, 3, 3, 3), (2 * z_dim,)
    return s


def make_state_dict(seed: int, **cfg) -> Dict[str, Tensor]:
    sd = {}
    for idx, (name, shape) in enumerate(param_shapes(**cfg).items()):
        g = torch.Generator().manual_seed(seed * 5297 + idx)
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
