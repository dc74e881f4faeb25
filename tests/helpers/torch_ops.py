"""TEST INFRASTRUCTURE: a torch-CPU stand-in for the `yume_b200.ops` entry points the engines call (DiT, the three VAE decoders, the two
VAE encoders), with the same argument meaning and buffer layouts as the C ABI (include/yume_b200.h). It lets the CPU suite drive the HOST
logic of the engines (weight re-packing, packer, token streams, folded normalisations, frame bookkeeping, tile plans, Ulysses layouts)
against the reference-generated fixtures without a GPU. Tests monkeypatch it in; it is never imported by the package, and the product
path has no CPU fallback (a missing library or a CPU tensor is an error there)."""
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


# ------------------------------------------------------------------------------------------------------------
# DiT engine (yume_b200/dit.py) stand-ins
# ------------------------------------------------------------------------------------------------------------
def _gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def gemm(a, w, bias, out, epilogue=YB_EPI_BF16, gate=None, tok_idx=None, block_n=0, n_split=0, split_stride=0, a_split=0,
         a_split_stride=0, shape=None, res=None, cta_pair=0, split_k=None):
    """Stand-in of yb_gemm_bf16: every epilogue, the gate table + token index of the
    adaLN gate, the Ulysses layouts (n_split: column block j of the output written to chunk j; a_split: A given as K chunks)."""
    if a_split:                                              # A[t, k] = a[k // a_split, t, k % a_split]
        M, K = shape
        A = a.reshape(-1)[: (K // a_split) * a_split_stride].view(K // a_split, a_split_stride)[:, : M * a_split]
        A = A.reshape(K // a_split, M, a_split).permute(1, 0, 2).reshape(M, K).float()
    else:
        A = a.float()
        M = A.shape[0]
    y = A @ w.float().t()
    if bias is not None:
        y = y + bias
    if epilogue == YB_EPI_GELU_TANH:
        y = _gelu_tanh(y)
    elif epilogue == YB_EPI_GELU_ERF:
        y = F.gelu(y)
    elif epilogue == YB_EPI_RES_BF16:
        y = y + res.float()
    if epilogue == YB_EPI_GATE_RES:
        if gate is not None:
            rows = tok_idx.long() if tok_idx is not None else torch.zeros(M, dtype=torch.long)
            y = y * gate[rows]
        out.add_(y)
        return out
    if n_split:
        N = w.shape[0]
        dst = out.reshape(-1)[: (N // n_split) * split_stride].view(N // n_split, split_stride)[:, : M * n_split]
        dst.view(N // n_split, M, n_split).copy_(y.view(M, N // n_split, n_split).permute(1, 0, 2).to(out.dtype))
        return out
    out.copy_(y.to(out.dtype))
    return out


YB_EPI_GELU_BF16, YB_EPI_GELU_ERF_BF16 = YB_EPI_GELU_TANH, YB_EPI_GELU_ERF


def ln_modulate(x, out, scale, shift, tok_idx=None, weight=None, bias=None, eps=1e-6):
    y = F.layer_norm(x.float(), (x.shape[1],), None, None, eps)
    if weight is not None:
        y = y * weight + bias
    if scale is not None:
        rows = tok_idx.long() if tok_idx is not None else torch.zeros(x.shape[0], dtype=torch.long)
        sc = scale[rows] if scale.dim() == 2 else scale
        sh = shift[rows] if shift.dim() == 2 else shift
        y = y * (1 + sc) + sh
    out.copy_(y.to(out.dtype))
    return out


def _norm_rope_rows(x, weight, rope, head_dim, eps, rope_len):
    """x f32 [L, C] -> RMSNorm over C * weight, then 3-axis RoPE on adjacent pairs of every head for rows < rope_len."""
    L, C = x.shape
    y = x * torch.rsqrt(x.pow(2).mean(dim=1, keepdim=True) + eps) * weight
    if rope is not None:
        n = L if rope_len is None else min(rope_len, L)
        v = y[:n].view(n, C // head_dim, head_dim // 2, 2)
        cos, sin = rope[:n, None, :, 0], rope[:n, None, :, 1]
        rot = torch.stack([v[..., 0] * cos - v[..., 1] * sin, v[..., 0] * sin + v[..., 1] * cos], dim=-1)
        y = torch.cat([rot.reshape(n, C), y[n:]], dim=0)
    return y


def _pieces_view(qk, pieces):
    """The peer-major Ulysses send buffer seen as [pieces, L, piece_cols]: logical column c of row t is element c % piece_cols of
    piece c // piece_cols, pieces `piece_stride` elements apart, rows qk.stride(0) apart (yb_rmsnorm_rope_pieces)."""
    L, C, pc, ps = pieces
    return torch.as_strided(qk, (C // pc, L, pc), (ps, qk.stride(0), 1), qk.storage_offset())


def rmsnorm_rope(qk, weight, rope, head_dim, eps=1e-6, rope_len=None, pieces=None):
    if pieces is not None:
        v = _pieces_view(qk, pieces)                                         # [P, L, pc]
        L, C = pieces[0], pieces[1]
        y = _norm_rope_rows(v.permute(1, 0, 2).reshape(L, C).float(), weight, rope, head_dim, eps, rope_len)
        v.copy_(y.view(L, C // pieces[2], pieces[2]).permute(1, 0, 2).to(qk.dtype))
        return qk
    qk.copy_(_norm_rope_rows(qk.float(), weight, rope, head_dim, eps, rope_len).to(qk.dtype))
    return qk


def qk_norm_rope(q, k, wq, wk, rope, head_dim, eps=1e-6, rope_len=None, pieces=None):
    rmsnorm_rope(q, wq, rope, head_dim, eps, rope_len, pieces)
    rmsnorm_rope(k, wk, rope, head_dim, eps, rope_len, pieces)


def attention(q, k, v, out, heads, scale=None, variant=0, accumulate=False, split=0):
    Lq, Lk, d = q.shape[0], k.shape[0], 128
    f = lambda t, L: t.float().reshape(L, heads, d).transpose(0, 1)[None]  # noqa: E731
    o = F.scaled_dot_product_attention(f(q, Lq), f(k, Lk), f(v, Lk), scale=scale)[0].transpose(0, 1).reshape(Lq, heads * d)
    if accumulate:
        o = o + out.float()
    out.copy_(o.to(out.dtype))
    return out


def linear_f32_small(x, w, bias, silu_in=False):
    y = (F.silu(x) if silu_in else x) @ w.t()
    return y + bias if bias is not None else y


def linear_f32(x, w, bias, out):
    out.copy_(x @ w.t() + (bias if bias is not None else 0))
    return out


def bcast_add(a, b):
    return a[:, None, :] + b[None, :, :]


def sinusoidal(t, dim):
    half = dim // 2
    pos = t.reshape(-1).double()
    s = torch.outer(pos, torch.pow(10000, -torch.arange(half).double().div(half)))
    return torch.cat([torch.cos(s), torch.sin(s)], dim=1).float()


def patchify(x, out, ph, pw):
    """x f32 [Cin, F, H, W] -> out bf16 [F*ceil(H/ph)*ceil(W/pw), >= Cin*ph*pw]: token order (f, hp, wp), column order (cin, i, j);
    rows / columns past H, W read as zero (convpadd)."""
    Cin, Fr, H, W = x.shape
    hp, wp = -(-H // ph), -(-W // pw)
    xp = F.pad(x.float(), (0, wp * pw - W, 0, hp * ph - H))
    t = xp.view(Cin, Fr, hp, ph, wp, pw).permute(1, 2, 4, 0, 3, 5).reshape(Fr * hp * wp, Cin * ph * pw)
    out[:, : Cin * ph * pw] = t.to(out.dtype)
    return out


def unpatchify(y, out, Fr, Hp, Wp, ph, pw):
    Cout = out.shape[0]
    out.copy_(y[: Fr * Hp * Wp, : ph * pw * Cout].reshape(Fr, Hp, Wp, ph, pw, Cout).permute(5, 0, 1, 3, 2, 4).reshape(Cout, Fr, Hp * ph, Wp * pw))
    return out
