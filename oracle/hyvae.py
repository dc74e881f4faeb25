"""CPU oracle for the HunyuanVideo causal 3D VAE *decode* path — TEST INFRASTRUCTURE ONLY.

Functional restatement (plain PyTorch, CPU) of `AutoencoderKLCausal3D.decode` and everything under it:
  /root/reference/hyvideo/vae/autoencoder_kl_causal_3d.py:297-359 (decode, blend_v/h/t), :417-463, :500-531 (tiling)
  /root/reference/hyvideo/vae/vae.py:131-291 (DecoderCausal3D)
  /root/reference/hyvideo/vae/unet_causal_3d_blocks.py:37-74 (mask, CausalConv3d), :129-182 (UpsampleCausal3D),
      :348-415 (ResnetBlockCausal3D), :615-628 (UNetMidBlockCausal3D.forward), :754-764 (UpDecoderBlockCausal3D)
State-dict keys are the reference module tree's (`decoder.*`, `post_quant_conv.*`).

Third-party arithmetic not in the tree: the mid-block attention is `diffusers.models.attention_processor.Attention`
(pinned diffusers==0.32.0, requirements.txt:27) constructed at unet_causal_3d_blocks.py:580-592. Restated here from
its documented behaviour for those constructor arguments (GroupNorm(32) on [B,C,L] -> to_q/to_k/to_v Linear ->
single-head attention with the additive frame-causal mask, fp32 softmax -> to_out[0] Linear -> + residual ->
/ rescale_output_factor(=1)). No reference test pins that boundary: parity there is UNPINNED. The VAE
config JSON is not in the tree either; the upstream "884-16c-hy" values are pinned in CONFIG_884_16C below
(SURVEY.md §8c). The rest of this file is pinned by tests/golden/hyvae_tiny.pt, produced by tools/make_golden_vae.py
from the reference's own code with the same Attention restatement.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

CONFIG_884_16C = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512),
                      layers_per_block=2, norm_num_groups=32, act_fn="silu", sample_size=256, sample_tsize=64,
                      scaling_factor=0.476986, time_compression_ratio=4, spatial_compression_ratio=8,
                      mid_block_add_attention=True)


def causal_conv3d(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """CausalConv3d.forward (unet_causal_3d_blocks.py:72-74): replicate pad (W 1,1; H 1,1; T k-1,0) then Conv3d."""
    k = w.shape[2]
    if k > 1:
        x = F.pad(x, (k // 2, k // 2, k // 2, k // 2, k - 1, 0), mode="replicate")
    return F.conv3d(x, w, b)


def upsample_causal(x: Tensor, factor: Sequence[int]) -> Tensor:
    """UpsampleCausal3D.forward interpolate branch (:144-174): first frame is only upsampled spatially."""
    first, other = x.split((1, x.shape[2] - 1), dim=2)
    first = F.interpolate(first.squeeze(2), scale_factor=tuple(factor[1:]), mode="nearest").unsqueeze(2)
    if x.shape[2] > 1:
        other = F.interpolate(other, scale_factor=tuple(factor), mode="nearest")
        return torch.cat((first, other), dim=2)
    return first


def causal_attention_mask(n_frame: int, n_hw: int) -> Tensor:
    """prepare_causal_attention_mask (:37-45): token i sees every token of frames <= frame(i)."""
    frame = torch.arange(n_frame * n_hw) // n_hw
    mask = torch.full((n_frame * n_hw, n_frame * n_hw), float("-inf"))
    mask[frame[:, None] >= frame[None, :]] = 0
    return mask


class HyVaeOracle:
    def __init__(self, sd: Dict[str, Tensor], block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, sample_size=256, sample_tsize=64, time_compression_ratio=4,
                 spatial_compression_ratio=8, tile_overlap_factor=0.25, **_):
        self.sd, self.boc, self.lpb, self.groups = sd, tuple(block_out_channels), layers_per_block, norm_num_groups
        # tiling parameters, autoencoder_kl_causal_3d.py:122-132
        self.tile_sample_min_tsize = sample_tsize
        self.tile_latent_min_tsize = sample_tsize // time_compression_ratio
        self.tile_sample_min_size = sample_size
        self.tile_latent_min_size = int(sample_size / (2 ** (len(self.boc) - 1)))
        self.tile_overlap_factor = tile_overlap_factor
        self.use_spatial_tiling = self.use_temporal_tiling = False
        # which up blocks upsample what (vae.py:176-195)
        n = len(self.boc)
        self.up_factors: List[tuple] = []
        for i in range(n):
            sp = i < 3                                      # log2(spatial_compression_ratio) = 3
            tm = (i >= n - 1 - 2) and (i != n - 1)           # log2(time_compression_ratio) = 2
            self.up_factors.append(((2 if tm else 1), (2 if sp else 1), (2 if sp else 1)) if (sp or tm) else None)

    def enable_tiling(self, on: bool = True):
        self.use_spatial_tiling = self.use_temporal_tiling = on

    # ---- blocks --------------------------------------------------------------------------------------------
    def _gn(self, p: str, x: Tensor) -> Tensor:
        return F.group_norm(x, self.groups, self.sd[p + ".weight"], self.sd[p + ".bias"], eps=1e-6)

    def _conv(self, p: str, x: Tensor) -> Tensor:
        return causal_conv3d(x, self.sd[p + ".conv.weight"], self.sd[p + ".conv.bias"])

    def resnet(self, p: str, x: Tensor) -> Tensor:
        """ResnetBlockCausal3D.forward (:348-415) with temb=None, no up/down, output_scale_factor 1."""
        h = self._conv(p + ".conv1", F.silu(self._gn(p + ".norm1", x)))
        h = self._conv(p + ".conv2", F.silu(self._gn(p + ".norm2", h)))
        if (p + ".conv_shortcut.conv.weight") in self.sd:
            x = self._conv(p + ".conv_shortcut", x)
        return x + h

    def mid_attention(self, p: str, x: Tensor) -> Tensor:
        """UNetMidBlockCausal3D.forward attention step (:617-626) + the diffusers Attention restatement (see header)."""
        B, C, T, H, W = x.shape
        hs = x.permute(0, 2, 3, 4, 1).reshape(B, T * H * W, C)
        residual = hs
        hn = F.group_norm(hs.transpose(1, 2), self.groups, self.sd[p + ".group_norm.weight"],
                          self.sd[p + ".group_norm.bias"], eps=1e-6).transpose(1, 2)
        q = F.linear(hn, self.sd[p + ".to_q.weight"], self.sd[p + ".to_q.bias"])
        k = F.linear(hn, self.sd[p + ".to_k.weight"], self.sd[p + ".to_k.bias"])
        v = F.linear(hn, self.sd[p + ".to_v.weight"], self.sd[p + ".to_v.bias"])
        scores = torch.baddbmm(causal_attention_mask(T, H * W).to(q.dtype).expand(B, -1, -1), q, k.transpose(1, 2),
                               beta=1, alpha=C ** -0.5)
        probs = scores.float().softmax(dim=-1).to(q.dtype)   # upcast_softmax=True
        o = F.linear(torch.bmm(probs, v), self.sd[p + ".to_out.0.weight"], self.sd[p + ".to_out.0.bias"])
        o = o + residual
        return o.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)

    def decoder(self, z: Tensor) -> Tensor:
        """DecoderCausal3D.forward (vae.py:227-291)."""
        x = self._conv("decoder.conv_in", z)
        x = self.resnet("decoder.mid_block.resnets.0", x)
        x = self.mid_attention("decoder.mid_block.attentions.0", x)
        x = self.resnet("decoder.mid_block.resnets.1", x)
        for i in range(len(self.boc)):
            for j in range(self.lpb + 1):
                x = self.resnet(f"decoder.up_blocks.{i}.resnets.{j}", x)
            if self.up_factors[i] is not None:
                x = upsample_causal(x, self.up_factors[i])
                x = self._conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
        x = F.silu(self._gn("decoder.conv_norm_out", x))
        return self._conv("decoder.conv_out", x)

    def _decode_tile(self, z: Tensor) -> Tensor:
        z = F.conv3d(z, self.sd["post_quant_conv.weight"], self.sd["post_quant_conv.bias"])
        return self.decoder(z)

    # ---- tiling (autoencoder_kl_causal_3d.py:343-359, 417-463, 500-531) ---------------------------------------
    @staticmethod
    def _blend(a: Tensor, b: Tensor, extent: int, dim: int) -> Tensor:
        extent = min(a.shape[dim], b.shape[dim], extent)
        for y in range(extent):
            ia = [slice(None)] * 5
            ib = [slice(None)] * 5
            ia[dim], ib[dim] = a.shape[dim] - extent + y, y
            b[tuple(ib)] = a[tuple(ia)] * (1 - y / extent) + b[tuple(ib)] * (y / extent)
        return b

    def spatial_tiled_decode(self, z: Tensor) -> Tensor:
        overlap = int(self.tile_latent_min_size * (1 - self.tile_overlap_factor))
        blend_extent = int(self.tile_sample_min_size * self.tile_overlap_factor)
        row_limit = self.tile_sample_min_size - blend_extent
        rows = []
        for i in range(0, z.shape[-2], overlap):
            rows.append([self._decode_tile(z[:, :, :, i:i + self.tile_latent_min_size, j:j + self.tile_latent_min_size])
                         for j in range(0, z.shape[-1], overlap)])
        result_rows = []
        for i, row in enumerate(rows):
            result_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self._blend(rows[i - 1][j], tile, blend_extent, 3)
                if j > 0:
                    tile = self._blend(row[j - 1], tile, blend_extent, 4)
                result_row.append(tile[:, :, :, :row_limit, :row_limit])
            result_rows.append(torch.cat(result_row, dim=-1))
        return torch.cat(result_rows, dim=-2)

    def temporal_tiled_decode(self, z: Tensor) -> Tensor:
        overlap = int(self.tile_latent_min_tsize * (1 - self.tile_overlap_factor))
        blend_extent = int(self.tile_sample_min_tsize * self.tile_overlap_factor)
        t_limit = self.tile_sample_min_tsize - blend_extent
        row = []
        for i in range(0, z.shape[2], overlap):
            tile = z[:, :, i:i + self.tile_latent_min_tsize + 1]
            if self.use_spatial_tiling and (tile.shape[-1] > self.tile_latent_min_size or tile.shape[-2] > self.tile_latent_min_size):
                dec = self.spatial_tiled_decode(tile)
            else:
                dec = self._decode_tile(tile)
            if i > 0:
                dec = dec[:, :, 1:]
            row.append(dec)
        out = []
        for i, tile in enumerate(row):
            if i > 0:
                tile = self._blend(row[i - 1], tile, blend_extent, 2)
                out.append(tile[:, :, :t_limit])
            else:
                out.append(tile[:, :, :t_limit + 1])
        return torch.cat(out, dim=2)

    @torch.no_grad()
    def decode(self, z: Tensor) -> Tensor:
        """AutoencoderKLCausal3D.decode -> _decode (:297-341); z [1,16,T,H,W] -> [1,3,4(T-1)+1,8H,8W]."""
        assert z.dim() == 5
        if self.use_temporal_tiling and z.shape[2] > self.tile_latent_min_tsize:
            return self.temporal_tiled_decode(z)
        if self.use_spatial_tiling and (z.shape[-1] > self.tile_latent_min_size or z.shape[-2] > self.tile_latent_min_size):
            return self.spatial_tiled_decode(z)
        return self._decode_tile(z)


# ------------------------------------------------------------------------------------------------------------
# synthetic weights (seeded, bf16-representable) with the reference's state-dict keys
# ------------------------------------------------------------------------------------------------------------
def param_shapes(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=16, out_channels=3) -> Dict[str, tuple]:
    boc = list(block_out_channels)
    rev = boc[::-1]
    s: Dict[str, tuple] = {"post_quant_conv.weight": (latent_channels, latent_channels, 1, 1, 1), "post_quant_conv.bias": (latent_channels,)}

    def conv(p, ci, co, k=3):
        s[p + ".conv.weight"], s[p + ".conv.bias"] = (co, ci, k, k, k), (co,)

    def norm(p, c):
        s[p + ".weight"], s[p + ".bias"] = (c,), (c,)

    def resnet(p, ci, co):
        norm(p + ".norm1", ci); conv(p + ".conv1", ci, co); norm(p + ".norm2", co); conv(p + ".conv2", co, co)
        if ci != co:
            conv(p + ".conv_shortcut", ci, co, 1)

    c = rev[0]
    conv("decoder.conv_in", latent_channels, c)
    resnet("decoder.mid_block.resnets.0", c, c)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[f"{a}.{n}.weight"], s[f"{a}.{n}.bias"] = (c, c), (c,)
    resnet("decoder.mid_block.resnets.1", c, c)
    prev = c
    n = len(boc)
    for i in range(n):
        co = rev[i]
        for j in range(layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if i < 3 or ((i >= n - 3) and i != n - 1):
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co)
        prev = co
    norm("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", boc[0], out_channels)
    return s


def make_state_dict(seed: int, **cfg) -> Dict[str, Tensor]:
    sd = {}
    for idx, (name, shape) in enumerate(param_shapes(**cfg).items()):
        g = torch.Generator().manual_seed(seed * 7919 + idx)
        t = torch.randn(shape, generator=g)
        if name.endswith(".bias"):
            t = 0.05 * t
        elif len(shape) == 1:
            t = 1.0 + 0.1 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (1.5 / fan_in ** 0.5)
        sd[name] = t.to(torch.bfloat16).float()
    return sd
