"""World-size-2 `gloo` test (CPU) of the Ulysses host logic the engine relies on: the peer-major q|k|v weight
permutation (`sp_qkv_row_order`), the token shard arithmetic (`sp_shard`), and the two all-to-all layouts —
[P, Lp, q|k|v of heads/P] out, [P, Lp, heads/P*128] back, read K-split by the o-projection. The CUDA kernels are
replaced by plain torch CPU math *in this test only*; what is checked is that after both exchanges every rank
holds exactly the attention output rows of its own token shard in natural head order."""
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from yume_b200.dit import sp_qkv_row_order, sp_shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _attention(q, k, v, heads):
    Lq, Lk, d = q.shape[0], k.shape[0], q.shape[1] // heads
    f = lambda t, L: t.reshape(L, heads, d).transpose(0, 1)[None]  # noqa: E731
    o = F.scaled_dot_product_attention(f(q, Lq), f(k, Lk), f(v, Lk))[0]
    return o.transpose(0, 1).reshape(Lq, heads * d)


def _worker(rank, world, port, L, C, heads, errs):
    try:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        g = torch.Generator().manual_seed(0)
        h_full = torch.randn(L, C, generator=g, dtype=torch.float64)
        w_qkv = torch.randn(3 * C, C, generator=g, dtype=torch.float64) / C ** 0.5
        Lp, r0, nv = sp_shard(L, world, rank)
        h_loc = torch.zeros(Lp, C, dtype=torch.float64)
        h_loc[:nv] = h_full[r0:r0 + nv]
        Wh = C // world
        W3 = 3 * Wh
        # what the n_split GEMM epilogue writes: column block p (width W3) of the permuted output -> chunk p
        qkv_loc = h_loc @ w_qkv[sp_qkv_row_order(C, heads, world)].t()
        send = qkv_loc.view(Lp, world, W3).permute(1, 0, 2).contiguous()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        full = recv.view(world * Lp, W3)
        o = _attention(full[:, :Wh], full[:L, Wh:2 * Wh], full[:L, 2 * Wh:], heads // world)   # keys >= L masked
        att_send = o.reshape(world, Lp, Wh).contiguous()
        att_recv = torch.empty_like(att_send)
        dist.all_to_all_single(att_recv, att_send)
        a_loc = att_recv.permute(1, 0, 2).reshape(Lp, C)       # the K-split read: A[t, k] = att_recv[k // Wh, t, k % Wh]
        # single-process reference
        qkv = h_full @ w_qkv.t()
        ref = _attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads)
        if not torch.allclose(a_loc[:nv], ref[r0:r0 + nv], atol=1e-9):
            errs.put(f"rank {rank}: mismatch {float((a_loc[:nv] - ref[r0:r0 + nv]).abs().max())}")
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        errs.put(f"rank {rank}: {type(e).__name__}: {e}")


@pytest.mark.parametrize("L,C,heads", [(37, 64, 4), (48, 96, 6)])
def test_ulysses_layouts_world2(L, C, heads):
    ctx = mp.get_context("spawn")
    errs = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, L, C, heads, errs)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "worker crashed or timed out"
    msgs = []
    while not errs.empty():
        msgs.append(errs.get())
    assert not msgs, msgs


def test_shard_arithmetic():
    assert sp_shard(18480, 8, 0) == (2310, 0, 2310)
    assert sp_shard(18480, 8, 7) == (2310, 16170, 2310)
    assert sp_shard(37, 2, 1) == (19, 19, 18)
    assert sp_shard(9460, 8, 7) == (1183, 8281, 1179)
    assert sp_shard(5, 8, 6) == (1, 6, 0)
    idx = sp_qkv_row_order(3072, 24, 8)
    assert idx.numel() == 3 * 3072 and idx.unique().numel() == idx.numel()
    assert idx[:384].tolist() == list(range(0, 384)) and idx[384:768].tolist() == list(range(3072, 3072 + 384))
