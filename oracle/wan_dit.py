"""CPU oracle for the YUME denoise forward — TEST INFRASTRUCTURE ONLY.

A functional restatement (plain PyTorch on CPU, fp32 weights, bf16 attention inputs — the regime the reference
itself runs in on a CPU-only box) of the reference's `WanModel.forward` for both trees:

  * 5B  (Wan2.2-TI2V style):  /root/reference/wan23/modules/model.py
  * 14B (Wan2.1-I2V style):   /root/reference/wan/modules/model.py

Each function cites the reference lines it follows. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg may import this module; the product path (yume_b200/) never does.

Pinning: the reference ships no tests or golden vectors. The oracle is pinned instead against
outputs of the reference's own code imported in the authoring container (tools/make_golden.py ->
tests/golden/*.pt; checked by tests/test_oracle_golden.py). The third-party flash_attn kernel the reference calls
(attention.py:113-127, flash_attn==2.7.0.post2) is absent there; both the golden generator and this oracle use
the same SDPA restatement of its contract, so parity at that boundary is "unpinned" (stated in DESIGN.md).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------------------
# leaf functions
# ------------------------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim: int, position: Tensor) -> Tensor:
    """wan23/modules/model.py:14-24 — fp64 sinusoid, [cos | sin]."""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_params(max_seq_len: int, dim: int, theta: float = 10000) -> Tensor:
    """wan23/modules/model.py:27-35 — complex128 table [max_seq_len, dim/2]."""
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def rope_tables(head_dim: int):
    """The three per-axis tables (t, h, w): model.py:475-480 / :596-598."""
    d = head_dim
    return (rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6)))


def grid_freqs(tables, f: int, h: int, w: int, f0: int = 0) -> Tensor:
    """Per-token complex multipliers of a regular (f,h,w) grid: model.py:64-69 (grid path) and up_fre :933-940
    (f0 = temporal offset of the segment)."""
    t0, t1, t2 = tables
    return torch.cat([
        t0[f0:f0 + f].view(f, 1, 1, -1).expand(f, h, w, -1),
        t1[:h].view(1, h, 1, -1).expand(f, h, w, -1),
        t2[:w].view(1, 1, w, -1).expand(f, h, w, -1)
    ], dim=-1).reshape(f * h * w, 1, -1)


def rope_apply(x: Tensor, freqs_tok: Tensor) -> Tensor:
    """x [L, N, D] -> rotated fp32; tokens >= freqs_tok.shape[0] pass through (model.py:62-73, 102-106).
    Pairs are adjacent (2j, 2j+1) (view_as_complex of reshape(..., -1, 2))."""
    L, n, d = x.shape
    s = freqs_tok.shape[0]
    xc = torch.view_as_complex(x[:s].to(torch.float64).reshape(s, n, -1, 2))
    xr = torch.view_as_real(xc * freqs_tok).flatten(2)
    return torch.cat([xr, x[s:].to(torch.float64)]).float()


def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """WanRMSNorm.forward, model.py:129-137."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * weight


def layer_norm(x: Tensor, eps: float, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None) -> Tensor:
    """WanLayerNorm.forward, model.py:145-150."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps).type_as(x)


def flash_attention(q: Tensor, k: Tensor, v: Tensor, k_lens: Optional[Sequence[int]] = None) -> Tensor:
    """Contract of wan23/modules/attention.py:24-130 with SDPA standing in for flash_attn_varlen_func:
    q,k,v [B, L, N, D] cast to bf16 (`half`, :59-83), non-causal, scale 1/sqrt(D), keys >= k_lens[b] dropped,
    result returned in q's dtype (:130)."""
    out_dtype = q.dtype
    b, lq, lk = q.shape[0], q.shape[1], k.shape[1]
    qh, kh, vh = (t.to(torch.bfloat16).transpose(1, 2) for t in (q, k, v))
    mask = None
    if k_lens is not None and any(int(n) < lk for n in k_lens):   # (all keys valid: no mask, like the varlen kernel)
        mask = torch.zeros(b, 1, 1, lk, dtype=torch.bool, device=q.device)
        for i, n in enumerate(k_lens):
            mask[i, ..., :int(n)] = True
    o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask)
    return o.transpose(1, 2).contiguous().type(out_dtype)


def convpadd(t: Tensor, pad_num: int) -> Tensor:
    """model.py:918-931 — zero-pad H, W (bottom/right) of [B,C,F,H,W] to a multiple of pad_num."""
    if t.dim() == 4:
        t = t.unsqueeze(2)
    b, c, f, h, w = t.shape
    ph = (pad_num - h % pad_num) % pad_num
    pw = (pad_num - w % pad_num) % pad_num
    return F.pad(t, (0, pw, 0, ph))


# ------------------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------------------
class WanOracle:
    """Functional WanModel. `sd` uses the reference's state-dict keys; `variant` is '5b' or '14b'."""

    def __init__(self, sd: Dict[str, Tensor], variant: str, dim: int, ffn_dim: int, num_heads: int, num_layers: int,
                 in_dim: int, out_dim: int, text_len: int = 512, freq_dim: int = 256, patch_size=(1, 2, 2),
                 eps: float = 1e-6):
        assert variant in ("5b", "14b")
        self.sd, self.variant = sd, variant
        self.dim, self.ffn_dim, self.num_heads, self.num_layers = dim, ffn_dim, num_heads, num_layers
        self.in_dim, self.out_dim, self.text_len, self.freq_dim = in_dim, out_dim, text_len, freq_dim
        self.patch_size, self.eps = tuple(patch_size), eps
        self.d = dim // num_heads
        self.tables = rope_tables(self.d)
        self.trace: Dict[str, Tensor] = {}  # optional per-stage dumps (block 0) for stage-by-stage parity

    # ---- small helpers -------------------------------------------------------------------------------------
    def _lin(self, name: str, x: Tensor) -> Tensor:
        return F.linear(x, self.sd[name + ".weight"], self.sd.get(name + ".bias"))

    def _embed(self, name: str, u: Tensor) -> Tensor:
        """Conv3d with kernel == stride (patch_embedding*, model.py:453-454, 486-494)."""
        w = self.sd[name + ".weight"]
        return F.conv3d(u, w, self.sd.get(name + ".bias"), stride=w.shape[2:])

    # ---- attention -----------------------------------------------------------------------------------------
    def self_attn(self, p: str, x: Tensor, freqs_tok: Tensor, k_len: Optional[int] = None) -> Tensor:
        """WanSelfAttention.forward, model.py:178-207. k_lens: the 5B tree overrides seq_lens with x.shape[1] (:846-851)
        so every row is a key; the 14B tree passes the ACTUAL token count (wan/modules/model.py:311-314, 916) so the
        zero-padded rows of a regular-grid input with seq_len > F*H*W are masked (k_len)."""
        b, s, n, d = x.shape[0], x.shape[1], self.num_heads, self.d
        q = rms_norm(self._lin(p + ".q", x), self.sd[p + ".norm_q.weight"], self.eps).view(b, s, n, d)
        k = rms_norm(self._lin(p + ".k", x), self.sd[p + ".norm_k.weight"], self.eps).view(b, s, n, d)
        v = self._lin(p + ".v", x).view(b, s, n, d)
        q = torch.stack([rope_apply(q[i], freqs_tok) for i in range(b)])
        k = torch.stack([rope_apply(k[i], freqs_tok) for i in range(b)])
        if p.startswith("blocks.0."):
            self.trace["q_rope"], self.trace["k_rope"], self.trace["v"] = q, k, v
        o = flash_attention(q, k, v, k_lens=[s if k_len is None else k_len] * b)
        if p.startswith("blocks.0."):
            self.trace["attn_out"] = o
        return self._lin(p + ".o", o.flatten(2))

    def cross_attn(self, p: str, x: Tensor, context: Tensor) -> Tensor:
        """5B: WanCrossAttention.forward model.py:212-232; 14B: WanI2VCrossAttention wan/modules/model.py:363-389.
        context_lens is None on every path (:815) — padded text rows are NOT masked."""
        b, n, d = x.size(0), self.num_heads, self.d
        q = rms_norm(self._lin(p + ".q", x), self.sd[p + ".norm_q.weight"], self.eps).view(b, -1, n, d)
        if self.variant == "14b":
            ctx_img, ctx = context[:, :257], context[:, 257:]
        else:
            ctx_img, ctx = None, context
        k = rms_norm(self._lin(p + ".k", ctx), self.sd[p + ".norm_k.weight"], self.eps).view(b, -1, n, d)
        v = self._lin(p + ".v", ctx).view(b, -1, n, d)
        o = flash_attention(q, k, v).flatten(2)
        if ctx_img is not None:
            k_img = rms_norm(self._lin(p + ".k_img", ctx_img), self.sd[p + ".norm_k_img.weight"], self.eps)
            v_img = self._lin(p + ".v_img", ctx_img)
            o = o + flash_attention(q, k_img.view(b, -1, n, d), v_img.view(b, -1, n, d)).flatten(2)
        return self._lin(p + ".o", o)

    # ---- block ---------------------------------------------------------------------------------------------
    def block(self, i: int, x: Tensor, e0: Tensor, freqs_tok: Tensor, context: Tensor, k_len: Optional[int] = None) -> Tensor:
        """WanAttentionBlock.forward: 5B model.py:272-316 (e0 [B,L,6,C]); 14B wan/modules/model.py:444-496
        (e0 [B,6,C])."""
        p = f"blocks.{i}"
        mod = self.sd[p + ".modulation"]  # [1, 6, C]
        if self.variant == "5b":
            e = [t.squeeze(2) for t in (mod.unsqueeze(0) + e0).chunk(6, dim=2)]  # 6 x [B, L, C]
        else:
            e = list((mod + e0).chunk(6, dim=1))  # 6 x [B, 1, C]
        h = layer_norm(x, self.eps).float() * (1 + e[1]) + e[0]
        if i == 0:
            self.trace["h_norm1"] = h
        y = self.self_attn(p + ".self_attn", h, freqs_tok, k_len)
        x = x + y * e[2]
        xn = layer_norm(x, self.eps, self.sd[p + ".norm3.weight"], self.sd[p + ".norm3.bias"])
        x = x + self.cross_attn(p + ".cross_attn", xn, context)
        h = layer_norm(x, self.eps).float() * (1 + e[4]) + e[3]
        y = self._lin(p + ".ffn.2", F.gelu(self._lin(p + ".ffn.0", h), approximate="tanh"))
        x = x + y * e[5]
        if i == 0:
            self.trace["block0_out"] = x
        return x

    # ---- FramePack packer ----------------------------------------------------------------------------------
    def _segments(self, hist: int, branch_hist: int):
        """History segmentation, model.py:599-718. Returns a list of
        (frame slice of u1, embedder name, convpadd multiple, pre_2x_f) in token order. `branch_hist` is what the
        branch conditions test: f_num - latent_frame_zero (5B) or f_num - 9 (14B, wan/modules/model.py:779)."""
        H = hist
        one = lambda i: slice(i, i + 1) if i != -1 else slice(H - 1, H)  # noqa: E731
        neg = lambda i: slice(H + i, H + i + 1)  # noqa: E731  single frame at negative index i
        P1, P2, P4, P8, P16 = ("patch_embedding", 0), ("patch_embedding_2x", 4), ("patch_embedding_4x", 8), \
            ("patch_embedding_8x", 16), ("patch_embedding_16x", 32)
        if branch_hist <= 6:
            mid = neg(-1) if H - 2 <= 0 else slice(1, H - 1)
            return [(one(0), *P1, False), (mid, *P2, False), (neg(-1), *P1, False)]
        if branch_hist <= 22:
            mid = neg(-5) if H - 6 <= 0 else slice(1, H - 5)
            return [(one(0), *P1, False), (mid, *P4, False), (slice(H - 5, H - 3), *P2, False),
                    (slice(H - 3, H), *P1, False)]
        if branch_hist <= 86:
            mid = neg(-21) if H - 22 <= 0 else slice(1, H - 21)
            return [(one(0), *P1, False), (mid, *P8, False), (slice(H - 21, H - 5), *P4, False),
                    (slice(H - 5, H - 3), *P2, False), (slice(H - 3, H), *P1, False)]
        if branch_hist <= 342:
            mid = neg(-85) if H - 86 <= 0 else slice(1, H - 85)
            return [(one(0), *P2, False), (mid, *P16, False), (slice(H - 85, H - 21), *P8, False),
                    (slice(H - 21, H - 5), *P4, False), (slice(H - 5, H - 3), *P2, False),
                    (slice(H - 3, H), *P1, False)]
        if branch_hist <= 1366:
            mid = neg(-341) if H - 342 <= 0 else slice(1, H - 341)
            return [(one(0), *P2, False), (mid, *P16, True), (slice(H - 341, H - 85), *P16, False),
                    (slice(H - 85, H - 21), *P8, False), (slice(H - 21, H - 5), *P4, False),
                    (slice(H - 5, H - 3), *P2, False), (slice(H - 3, H), *P1, False)]
        raise UnboundLocalError("freqs_i: history longer than 1366 latent frames has no branch in the reference")

    def pack(self, u: Tensor, latent_frame_zero: int):
        """FramePack packer for one sample u [C,F,H,W] (model.py:591-729). Returns tokens [1,L,C], freqs [L,1,d/2],
        seq_lens1 (history token count), grid of the new frames."""
        u = u.unsqueeze(0)
        f_num = u.shape[2]
        u1, u2 = u[:, :, :-latent_frame_zero], u[:, :, -latent_frame_zero:]
        hist = u1.shape[2]
        branch_hist = f_num - (latent_frame_zero if self.variant == "5b" else 9)
        toks, freqs, f_z = [], [], 0
        for sl, name, padm, pre in self._segments(hist, branch_hist):
            seg = u1[:, :, sl]
            if pre:  # model.py:696-698: 2x_f (learned 4x4 stride-4 conv, in->in) before the 16x embedder
                seg = self._embed("patch_embedding_2x_f", convpadd(seg, 4))
            if padm:
                seg = convpadd(seg, padm)
            emb = self._embed(name, seg)
            _, _, f1, h1, w1 = emb.shape
            freqs.append(grid_freqs(self.tables, f1, h1, w1, f_z))
            f_z += f1
            toks.append(emb.flatten(2).transpose(1, 2))
        e2 = self._embed("patch_embedding", u2)
        freqs.append(grid_freqs(self.tables, *e2.shape[2:], f_z))
        seq_lens1 = sum(t.shape[1] for t in toks)
        grid = tuple(e2.shape[2:])
        toks.append(e2.flatten(2).transpose(1, 2))
        return torch.cat(toks, dim=1), torch.cat(freqs, dim=0), seq_lens1, grid

    # ---- embeddings ----------------------------------------------------------------------------------------
    def time_embed(self, t_flat: Tensor):
        """model.py:805-812: e = time_embedding(sinusoid(t)), e0 = time_projection(e); fp32."""
        e = sinusoidal_embedding_1d(self.freq_dim, t_flat).float()
        e = self._lin("time_embedding.2", F.silu(self._lin("time_embedding.0", e)))
        e0 = self._lin("time_projection.1", F.silu(e))
        return e, e0

    def text_embed(self, context: List[Tensor]) -> Tensor:
        """model.py:816-821: pad each to text_len rows with zeros, Linear-GELU(tanh)-Linear."""
        ctx = torch.stack([torch.cat([u, u.new_zeros(self.text_len - u.size(0), u.size(1))]) for u in context])
        return self._lin("text_embedding.2", F.gelu(self._lin("text_embedding.0", ctx.float()), approximate="tanh"))

    def img_embed(self, clip_fea: Tensor) -> Tensor:
        """MLPProj, wan/modules/model.py:529-541 (LayerNorm eps default 1e-5, exact GELU)."""
        x = F.layer_norm(clip_fea.float(), (clip_fea.shape[-1],), self.sd["img_emb.proj.0.weight"],
                         self.sd["img_emb.proj.0.bias"])
        x = F.gelu(self._lin("img_emb.proj.1", x))
        x = self._lin("img_emb.proj.3", x)
        return F.layer_norm(x, (x.shape[-1],), self.sd["img_emb.proj.4.weight"], self.sd["img_emb.proj.4.bias"])

    def head(self, x: Tensor, e: Tensor) -> Tensor:
        """Head.forward: 5B model.py:336-348 (e [B,L,C]); 14B wan/modules/model.py:516-526 (e [B,C])."""
        mod = self.sd["head.modulation"]  # [1, 2, C]
        if self.variant == "5b":
            sh, sc = [t.squeeze(2) for t in (mod.unsqueeze(0) + e.unsqueeze(2)).chunk(2, dim=2)]
        else:
            sh, sc = (mod + e.unsqueeze(1)).chunk(2, dim=1)
        return self._lin("head.head", layer_norm(x, self.eps) * (1 + sc) + sh)

    def unpatchify(self, x: Tensor, grid) -> Tensor:
        """model.py:867-890."""
        c = self.out_dim
        u = x[:math.prod(grid)].view(*grid, *self.patch_size, c)
        u = torch.einsum("fhwpqrc->cfphqwr", u)
        return u.reshape(c, *[i * j for i, j in zip(grid, self.patch_size)])

    # ---- forward -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: List[Tensor], t: Tensor, context: List[Tensor], seq_len: int, y: Optional[List[Tensor]] = None,
                clip_fea: Optional[Tensor] = None, latent_frame_zero: Optional[int] = None, flag: bool = True,
                rand_num_img: Optional[float] = None) -> Tensor:
        """WanModel.forward: 5B model.py:547-865 (`flag` selects FramePack); 14B wan/modules/model.py:723-1013
        (`rand_num_img >= 0.4` selects FramePack, `< 0.4` the regular grid). Batch of one sample (as on every
        Yume path). Returns the fp32 output [C_out, F_new, H, W]."""
        assert len(x) == 1, "the Yume paths run batch 1 (rope_apply packed path only handles x[0], model.py:102)"
        if latent_frame_zero is None:
            latent_frame_zero = 8 if self.variant == "5b" else 9
        if y is not None:
            x = [torch.cat([u, v], dim=0) for u, v in zip(x, y)]
        if self.variant == "14b":
            assert clip_fea is not None and y is not None  # wan/modules/model.py:760-761
            assert rand_num_img is not None, "rand_num_img=None hits the packed RoPE path with a grid table (Appendix A)"
            packed = rand_num_img >= 0.4
        else:
            packed = bool(flag)

        k_len = None
        if packed:
            tok, freqs_tok, seq_lens1, grid = self.pack(x[0].float(), latent_frame_zero)
            L = tok.shape[1]
            if self.variant == "5b":  # per-token t: history tokens t[0], new tokens t[-1] (model.py:730-737)
                ts = t.squeeze()
                t_tok = torch.cat([ts[0:1].new_ones(seq_lens1) * ts[0], ts[-1:].new_ones(L - seq_lens1) * ts[-1]])
        else:
            emb = self._embed("patch_embedding", x[0].float().unsqueeze(0))
            grid = tuple(emb.shape[2:])
            tok = emb.flatten(2).transpose(1, 2)
            assert tok.shape[1] <= seq_len  # model.py:755
            tok = torch.cat([tok, tok.new_zeros(1, seq_len - tok.size(1), tok.size(2))], dim=1)
            L = tok.shape[1]
            freqs_tok = grid_freqs(self.tables, *grid)
            seq_lens1 = 0
            if self.variant == "14b":   # seq_lens = real token count -> k_lens (wan/modules/model.py:916, 311-314)
                k_len = math.prod(grid)
            if self.variant == "5b":
                t_tok = (t.expand(t.size(0), seq_len) if t.dim() == 1 else t).flatten()  # model.py:803-804

        if self.variant == "5b":
            e, e0 = self.time_embed(t_tok)
            e = e.unflatten(0, (1, L))
            e0 = e0.unflatten(0, (1, L)).unflatten(2, (6, self.dim))
        else:
            e, e0 = self.time_embed(t)  # wan/modules/model.py:924-928 ([B] -> [B,C], [B,6,C])
            e0 = e0.unflatten(1, (6, self.dim))

        ctx = self.text_embed(context)
        if self.variant == "14b":
            ctx = torch.cat([self.img_embed(clip_fea), ctx], dim=1)  # :939-941
        self.trace["context"] = ctx
        self.trace["tokens"] = tok

        xs = tok
        for i in range(self.num_layers):
            xs = self.block(i, xs, e0, freqs_tok, ctx, k_len)
        out = self.head(xs, e)
        self.trace["head_out"] = out
        return self.unpatchify(out[0, seq_lens1:], grid).float()
