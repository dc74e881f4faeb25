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
    assert oob_zero_pad and tuple(x.shape[:3]) == (T, H, W)
    kt, kh, kw = taps
    Cp, co = x.shape[-1], w.shape[0]
    wt = w.float().view(co, kt, kh, kw, Cp).permute(0, 4, 1, 2, 3)
    xn = x.float().permute(3, 0, 1, 2)[None]
    if stride_hw > 1:
        xn = F.pad(xn, (0, 1, 0, 1, 0, 0))
    else:
        xn = F.pad(xn, (kw // 2, kw // 2, kh // 2, kh // 2, 0, 0))
    if stride_t == 1:
        xn = F.pad(xn, (0, 0, 0, 0, kt - 1, 0))
    y = F.conv3d(xn, wt, bias, stride=(stride_t, stride_hw, stride_hw))[0].permute(1, 2, 3, 0)     # [To, Ho, Wo, co]
    To, Ho, Wo = y.shape[:3]
    assert (To, Ho, Wo) == conv_out_dims(T, H, W, taps, stride_t, stride_hw)
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
    s = S[:L].clone()
    s[:, hw:] = float("-inf")
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
