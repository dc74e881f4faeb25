"""TEST INFRASTRUCTURE: a torch-CPU stand-in for the handful of `yume_b200.ops` entry points the VAE encoder engines call, with
the same argument meaning and buffer layouts as the C ABI (include/yume_b200.h). It lets the CPU suite drive the HOST logic of
yume_b200/vae_enc.py (weight re-packing, folded normalisation, frame bookkeeping, strides, shortcut wiring) against the
reference-generated fixtures without a GPU. Never imported by the package; the product path has no CPU fallback."""
import torch
import torch.nn.functional as F

YB_EPI_BF16, YB_EPI_GELU_TANH, YB_EPI_F32, YB_EPI_GATE_RES, YB_EPI_GELU_ERF, YB_EPI_RES_BF16 = 0, 1, 2, 3, 4, 5
_BF = torch.bfloat16


def conv_out_dims(T, H, W, taps=(3, 3, 3), stride_t=1, stride_hw=1):
    kt, kh, kw = taps
    To = (T - kt) // stride_t + 1 if stride_t > 1 else T
    Ho, Wo = ((H + 1 - kh) // stride_hw + 1, (W + 1 - kw) // stride_hw + 1) if stride_hw > 1 else (H, W)
    return To, Ho, Wo


def nchw_to_nhwc_bf16(x, out):
    out.zero_()
    out[:, :x.shape[0]] = x.t().to(_BF)
    return out


def nhwc_to_nchw_f32(x, out, clamp=None):
    out.copy_(x[:, :out.shape[0]].t())
    if clamp is not None:
        out.clamp_(*clamp)
    return out


def gemm(a, w, bias, out, epilogue=YB_EPI_BF16, res=None, **_):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if epilogue == YB_EPI_RES_BF16:
        y = y + res.float()
    out.copy_(y.to(out.dtype))
    return out


def conv3d_causal(x, w, bias, out, T, H, W, epilogue=YB_EPI_BF16, res=None, taps=(3, 3, 3), oob_zero_pad=False, out_t_mul=1,
                  out_t_add=0, fuse_w=0, cta_pair=None, stride_t=1, stride_hw=1):
    kt, kh, kw = taps
    Cp, co = x.shape[-1], w.shape[0]
    wt = w.float().view(co, kt, kh, kw, Cp).permute(0, 4, 1, 2, 3)
    xn = x.float().permute(3, 0, 1, 2)[None]
    if not oob_zero_pad:           # hyvideo: the input already carries the replicate padding [T+2, H+2, W+2, Cp]
        assert tuple(x.shape[:3]) == (T + kt - 1, H + kh - 1, W + kw - 1) and stride_t == 1 and stride_hw == 1
    elif stride_hw > 1:
        xn = F.pad(xn, (0, 1, 0, 1, 0, 0))
    else:
        xn = F.pad(xn, (kw // 2, kw // 2, kh // 2, kh // 2, 0, 0))
    if oob_zero_pad and stride_t == 1:
        xn = F.pad(xn, (0, 0, 0, 0, kt - 1, 0))
    y = F.conv3d(xn, wt, bias, stride=(stride_t, stride_hw, stride_hw))[0].permute(1, 2, 3, 0)     # [To, Ho, Wo, co]
    To, Ho, Wo = y.shape[:3]
    assert (To, Ho, Wo) == (conv_out_dims(T, H, W, taps, stride_t, stride_hw) if oob_zero_pad else (T, H, W))
    y = y.reshape(To, Ho * Wo, co)
    if epilogue == YB_EPI_RES_BF16:
        y = y + res.float().view(To, Ho * Wo, co)
    frames = out.view(-1, Ho * Wo, out.shape[-1])
    for t in range(To):
        frames[t * out_t_mul + out_t_add, :, :co] = y[t].to(out.dtype)
    return out


def vae_rms_act(x, dims, out, gamma, up=1, silu=True):
    T, Hs, Ws = dims
    C = x.shape[1]
    y = x.float()
    if gamma is not None:
        y = F.normalize(y, dim=1) * (C ** 0.5) * gamma
    if silu:
        y = F.silu(y)
    y = y.view(T, Hs, Ws, C)
    if up == 2:
        y = y.repeat_interleave(2, 1).repeat_interleave(2, 2)
    out.zero_()
    out[..., :C] = y.to(out.dtype)
    return out


def masked_softmax(S, P, L, hw):
    # frame-causal: query row i (frame i // hw) sees key j iff j < L and j // hw <= i // hw (with L == hw: keys < hw only)
    s = S[:L].clone()
    rows = torch.arange(L)[:, None] // hw
    cols = torch.arange(S.shape[1])[None, :]
    s[(cols // hw > rows) | (cols >= L)] = float("-inf")
    P[:L] = torch.softmax(s, dim=-1).to(P.dtype)
    return P


def vae_avgdown_add(main, x, dims, in_c, out_c, ft, fs):
    T, H, W = dims
    xn = x.float().view(T, H, W, in_c).permute(3, 0, 1, 2)[None]
    pad_t = (ft - T % ft) % ft
    xn = F.pad(xn, (0, 0, 0, 0, pad_t, 0))
    B, C, Tp, _, _ = xn.shape
    xn = xn.view(B, C, Tp // ft, ft, H // fs, fs, W // fs, fs).permute(0, 1, 3, 5, 7, 2, 4, 6).contiguous()
    xn = xn.view(B, out_c, C * ft * fs * fs // out_c, Tp // ft, H // fs, W // fs).mean(dim=2)[0]
    main.copy_((main.float() + xn.permute(1, 2, 3, 0).reshape(-1, out_c)).to(main.dtype))
    return main


def vae_patchify2_bf16(video, out):
    c, f, H, W = video.shape
    x = video.view(c, f, H // 2, 2, W // 2, 2).permute(0, 5, 3, 1, 2, 4).reshape(c * 4, f, H // 2, W // 2)   # (c r q) f h w
    out.zero_()
    out[:, :12] = x.permute(1, 2, 3, 0).reshape(-1, 12).to(out.dtype)
    return out


def vae_dupup_add(main, x, dims, in_c, out_c, ft, fs):
    """main [ft*T-(ft-1), H*fs, W*fs, out_c] += DupUp3D(x [T, H, W, in_c]) (first ft-1 duplicated frames dropped)."""
    from oracle import wan22vae
    T, H, W = dims
    xn = x.float().view(T, H, W, in_c).permute(3, 0, 1, 2)[None]
    up = wan22vae.Wan22VaeOracle.dup_up(xn, out_c, ft, fs)[0].permute(1, 2, 3, 0)          # [To, Ho, Wo, out_c]
    main.copy_((main.float().view(up.shape) + up).reshape(main.shape).to(main.dtype))
    return main


def vae_unpatchify2_clamp(y, out, T, H, W):
    v = y[:, :12].view(T, H, W, 12).permute(3, 0, 1, 2)[None]
    out.copy_(v.reshape(1, 3, 2, 2, T, H, W).permute(0, 1, 4, 5, 3, 6, 2).reshape(3, T, 2 * H, 2 * W).clamp(-1, 1))
    return out


def gn_stats(x, groups):
    N, C = x.shape
    v = x.double().view(N, groups, C // groups)
    return torch.stack([v.sum(dim=(0, 2)), (v * v).sum(dim=(0, 2))], dim=1)


def vae_pad_act(x, src_dims, out, pad, up=(1, 1, 1), stats=None, gamma=None, beta=None, groups=32, eps=1e-6, silu=False):
    """[GroupNorm from the (sum, sum of squares) table] -> [SiLU] -> nearest upsample (first frame only spatially) -> replicate pad
    (2 frames in front, 1 voxel around) into out [T(+2), H(+2), W(+2), Cp]; channel padding zero."""
    Ts, Hs, Ws = src_dims
    C = x.shape[1]
    y = x.float().view(Ts, Hs, Ws, C)
    if stats is not None:
        cnt = Ts * Hs * Ws * (C // groups)
        mean = (stats[:, 0] / cnt)
        var = stats[:, 1] / cnt - mean * mean
        a = torch.rsqrt(var.float() + eps).repeat_interleave(C // groups) * gamma
        b = beta - mean.float().repeat_interleave(C // groups) * a
        y = y * a + b
    if silu:
        y = F.silu(y)
    ft, fh, fw = up
    if ft == 2:
        idx = torch.tensor([0] + [1 + ((t - 1) >> 1) for t in range(1, 1 + 2 * (Ts - 1))])
        y = y[idx]
    y = y.repeat_interleave(fh, 1).repeat_interleave(fw, 2)
    if pad:
        y = torch.cat([y[:1], y[:1], y], 0)
        y = torch.cat([y[:, :1], y, y[:, -1:]], 1)
        y = torch.cat([y[:, :, :1], y, y[:, :, -1:]], 2)
    out.zero_()
    out[..., :C] = y.to(out.dtype)
    return out


def vae_assemble_tiles(tiles, th, tw, tlen, tf0, out, row_limit, blend_extent, t_limit, t_blend_extent):
    """The reference's own sequence (autoencoder_kl_causal_3d.py:417-463, 500-531) on raw tiles [C, frames, h, w]: in-place blend_v,
    blend_h, crop, cat per temporal window; first frame of later windows dropped; blend_t, crop, cat."""
    def blend(a, b, extent, dim):
        extent = min(a.shape[dim], b.shape[dim], extent)
        for y in range(extent):
            ia, ib = [slice(None)] * 4, [slice(None)] * 4
            ia[dim], ib[dim] = a.shape[dim] - extent + y, y
            b[tuple(ib)] = a[tuple(ia)] * (1 - y / extent) + b[tuple(ib)] * (y / extent)
        return b

    windows = []
    for ti, plane in enumerate(tiles):
        rows = [[t.clone() for t in line] for line in plane]
        if len(rows) == 1 and len(rows[0]) == 1:
            dec = rows[0][0]
        else:
            result_rows = []
            for i, row in enumerate(rows):
                result_row = []
                for j, tile in enumerate(row):
                    if i > 0:
                        tile = blend(rows[i - 1][j], tile, blend_extent, 2)
                    if j > 0:
                        tile = blend(row[j - 1], tile, blend_extent, 3)
                    result_row.append(tile[:, :, :row_limit, :row_limit])
                result_rows.append(torch.cat(result_row, dim=-1))
            dec = torch.cat(result_rows, dim=-2)
        windows.append(dec[:, 1:] if ti > 0 else dec)
    if len(windows) == 1:
        out.copy_(windows[0])
        return out
    pieces = []
    for i, tile in enumerate(windows):
        if i > 0:
            tile = blend(windows[i - 1], tile, t_blend_extent, 1)
            pieces.append(tile[:, :t_limit])
        else:
            pieces.append(tile[:, :t_limit + 1])
    out.copy_(torch.cat(pieces, dim=1))
    return out
