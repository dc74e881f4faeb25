#!/usr/bin/env python
"""bench.py — YUME denoise-step benchmark on B200 (driver contract: one JSON line on stdout from rank 0).

Workload at N=1 = BASELINE.json configs[1]: one Yume-5B-720P denoise step (WanModel.forward, wan23 tree) on an
81-frame 704x1280 latent [48,21,44,80] -> L = 18 480 tokens, random-init bf16-valued weights, synthetic Gaussian
latent and text context, t = 500. metric = latent frames produced per second by the denoise forward (`ms_per_step`
is the denoise-step latency BASELINE.json names).

  value      device-resident inputs (latent, t, context already in HBM), CUDA-event timed
  e2e        the same step through the reference-facing `WanModel.forward(x, t, context, seq_len, flag=False)`
             with pinned HOST inputs: H2D copies and the D2H read of the result are inside the timed region
  roofline   the dominant kernel (self-attention, attention_kernel<true>) timed live with CUDA events on the
             launching stream inside the timed steps: achieved = 4*L^2*C FLOP / mean launch time
  cpu_baseline / --impl reference: the oracle port of the reference forward (oracle/wan_dit.py) on the host cores,
             on a bounded sample (one WanAttentionBlock of the same workload, extrapolated by FLOPs).
N > 1: Ulysses sequence parallelism over NCCL (one process per GPU, torchrun) — strong scaling of the same step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG_5B = dict(variant="5b", dim=3072, ffn_dim=14336, num_heads=24, num_layers=30, in_dim=48, out_dim=48,
              text_len=512, text_dim=4096, freq_dim=256)
CFG_14B = dict(variant="14b", dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, in_dim=36, out_dim=16,
               text_len=512, text_dim=4096, freq_dim=256, clip_dim=1280)
LATENT = (48, 21, 44, 80)          # 81 frames @ 704x1280 through the Wan2.2 VAE (stride 4,16,16)
SEQ_LEN = 21 * 22 * 40             # 18 480 tokens
CTX_LEN = 512


def block_flops(L: int, C: int, F: int, S: int) -> float:
    """8LC^2 + 4L^2C + 4LC^2 + 4SC^2 + 4LSC + 4LCF."""
    return 8 * L * C * C + 4 * L * L * C + 4 * L * C * C + 4 * S * C * C + 4 * L * S * C + 4 * L * C * F


def synthetic_state_dict(cfg: dict, device, seed: int = 0):
    """Random-init weights with the reference's state-dict keys, generated on the device, bf16-representable
    (the reference zero-inits head.head — model.py:914 — which would make the output 0; we use N(0, 0.02))."""
    g = torch.Generator(device=device).manual_seed(seed)
    C, Fd, cin, cout = cfg["dim"], cfg["ffn_dim"], cfg["in_dim"], cfg["out_dim"]

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g, device=device) * std).to(torch.bfloat16).to(torch.float32)

    sd = {}
    for name, k in (("patch_embedding", 2), ("patch_embedding_2x", 4), ("patch_embedding_4x", 8),
                    ("patch_embedding_8x", 16), ("patch_embedding_16x", 32)):
        sd[name + ".weight"] = rn(C, cin, 1, k, k, std=1 / math.sqrt(cin * k * k))
        sd[name + ".bias"] = rn(C, std=0.02)
    sd["patch_embedding_2x_f.weight"] = rn(cin, cin, 1, 4, 4, std=1 / math.sqrt(cin * 16))
    sd["patch_embedding_2x_f.bias"] = rn(cin, std=0.02)
    sd["text_embedding.0.weight"], sd["text_embedding.0.bias"] = rn(C, cfg["text_dim"], std=1 / 64), rn(C, std=0.02)
    sd["text_embedding.2.weight"], sd["text_embedding.2.bias"] = rn(C, C, std=C ** -0.5), rn(C, std=0.02)
    sd["time_embedding.0.weight"], sd["time_embedding.0.bias"] = rn(C, cfg["freq_dim"], std=1 / 16), rn(C, std=0.02)
    sd["time_embedding.2.weight"], sd["time_embedding.2.bias"] = rn(C, C, std=C ** -0.5), rn(C, std=0.02)
    sd["time_projection.1.weight"], sd["time_projection.1.bias"] = rn(6 * C, C, std=C ** -0.5), rn(6 * C, std=0.02)
    img = cfg["variant"] == "14b"
    if img:
        cd = cfg["clip_dim"]
        sd["img_emb.proj.0.weight"], sd["img_emb.proj.0.bias"] = 1 + rn(cd, std=0.1), rn(cd, std=0.02)
        sd["img_emb.proj.1.weight"], sd["img_emb.proj.1.bias"] = rn(cd, cd, std=cd ** -0.5), rn(cd, std=0.02)
        sd["img_emb.proj.3.weight"], sd["img_emb.proj.3.bias"] = rn(C, cd, std=cd ** -0.5), rn(C, std=0.02)
        sd["img_emb.proj.4.weight"], sd["img_emb.proj.4.bias"] = 1 + rn(C, std=0.1), rn(C, std=0.02)
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}"
        for att in ("self_attn", "cross_attn"):
            for pr in ("q", "k", "v", "o") + (("k_img", "v_img") if (img and att == "cross_attn") else ()):
                sd[f"{p}.{att}.{pr}.weight"], sd[f"{p}.{att}.{pr}.bias"] = rn(C, C, std=C ** -0.5), rn(C, std=0.02)
            sd[f"{p}.{att}.norm_q.weight"] = 1 + rn(C, std=0.1)
            sd[f"{p}.{att}.norm_k.weight"] = 1 + rn(C, std=0.1)
            if img and att == "cross_attn":
                sd[f"{p}.{att}.norm_k_img.weight"] = 1 + rn(C, std=0.1)
        sd[f"{p}.norm3.weight"], sd[f"{p}.norm3.bias"] = 1 + rn(C, std=0.1), rn(C, std=0.02)
        sd[f"{p}.ffn.0.weight"], sd[f"{p}.ffn.0.bias"] = rn(Fd, C, std=C ** -0.5), rn(Fd, std=0.02)
        sd[f"{p}.ffn.2.weight"], sd[f"{p}.ffn.2.bias"] = rn(C, Fd, std=Fd ** -0.5), rn(C, std=0.02)
        sd[f"{p}.modulation"] = rn(1, 6, C, std=C ** -0.5)
    sd["head.head.weight"], sd["head.head.bias"] = rn(4 * cout, C, std=0.02), rn(4 * cout, std=0.02)
    sd["head.modulation"] = rn(1, 2, C, std=C ** -0.5)
    return sd


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm (the one place bench.py executes oracle/): bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(budget_s: float, reps: int, warmup: int):
    """Times the oracle port of WanAttentionBlock.forward (oracle/wan_dit.py <- wan23/modules/model.py:272-316) at
    the real 5B width on the host cores. Sample = one block at L' tokens (L' the largest of {18480, 9240, 4620, 2310,
    1155} whose (reps+warmup) runs fit in `budget_s`), extrapolated to the 30-block step by algorithmic FLOPs."""
    from oracle import synth
    from oracle.wan_dit import WanOracle, grid_freqs

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dict(synth.CFG_5B, num_layers=1)
    sd = synth.make_state_dict(cfg, 1, num_layers=1)
    m = WanOracle(sd, **synth.oracle_kwargs(cfg))
    C, F, S = cfg["dim"], cfg["ffn_dim"], cfg["text_len"]

    def run(L, grid):
        g = torch.Generator().manual_seed(0)
        x = torch.randn(1, L, C, generator=g)
        e = 0.5 * torch.randn(1, L, 6, C, generator=g)
        ctx = torch.randn(1, S, C, generator=g)
        fr = grid_freqs(m.tables, *grid)
        t0 = time.perf_counter()
        with torch.no_grad():
            m.block(0, x, e, fr, ctx)
        return time.perf_counter() - t0

    ladder = [(18480, (21, 22, 40)), (9240, (21, 22, 20)), (4620, (21, 11, 20)), (2310, (21, 11, 10)), (1155, (21, 11, 5))]
    run(*ladder[-1])                                           # first touch of weights / thread pool: discarded
    probe = run(*ladder[-1])
    per_flop = probe / block_flops(1155, C, F, S)
    choice = ladder[-1]
    for L, grid in ladder:
        if per_flop * block_flops(L, C, F, S) * (reps + warmup) <= budget_s:
            choice = (L, grid)
            break
    L, grid = choice
    for _ in range(max(0, warmup - 1)):
        run(L, grid)
    times = [run(L, grid) for _ in range(reps)]
    t_sample = sum(times) / len(times)
    step_flops = CFG_5B["num_layers"] * block_flops(SEQ_LEN, C, F, S)
    t_step = t_sample * step_flops / block_flops(L, C, F, S)
    return dict(t_sample_s=t_sample, t_step_s=t_step, cores=cores, L=L,
                sample=f"1 of 30 WanAttentionBlocks (oracle port, fp32 weights, bf16 SDPA) at L={L} of 18480 tokens, "
                       f"{reps} reps, extrapolated to the 30-block step by algorithmic FLOPs "
                       f"({step_flops / block_flops(L, C, F, S):.1f}x)")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_sample(budget_s=args.ref_budget, reps=args.steps, warmup=max(1, args.warmup))
    value = LATENT[1] / r["t_step_s"]
    line = {
        "impl": "reference", "metric": "latent_frames_per_sec", "value": value, "unit": "latent-frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["t_step_s"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (bf16 attention inputs)",
        "data": "synthetic",
        "config": {"workload": "Yume-5B-720P single denoise step, 81-frame 704x1280 latent [48,21,44,80], L=18480",
                   "timing": "time.perf_counter on host"},
        "cpu_baseline": {"value": value, "unit": "latent-frames/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": value, "unit": "latent-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------------------
def product_arm(args):
    import torch.distributed as dist

    from yume_b200 import ops
    from yume_b200.model import WanModel5B
    from yume_b200.utils import ClockSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    cfg = CFG_5B
    sd = synthetic_state_dict(cfg, dev, seed=0)
    with torch.device("meta"):
        model = WanModel5B(model_type="ti2v", text_len=cfg["text_len"], in_dim=cfg["in_dim"], dim=cfg["dim"],
                           ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"],
                           out_dim=cfg["out_dim"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"])
    model.install(dev, state_dict=sd)
    del sd
    eng = model._yb_engine
    if world > 1:
        eng.enable_sequence_parallel(dist.group.WORLD, transport=args.sp_transport)

    g = torch.Generator().manual_seed(1)
    x_host = torch.randn(*LATENT, generator=g).pin_memory()
    ctx_host = torch.randn(CTX_LEN, cfg["text_dim"], generator=g).to(torch.bfloat16).pin_memory()
    t_host = torch.tensor([500.0]).pin_memory()
    out_host = torch.empty(LATENT, dtype=torch.float32).pin_memory()
    x_dev, ctx_dev, t_dev = x_host.to(dev), ctx_host.to(dev), t_host.to(dev)

    def step_device():
        return eng.forward(x_dev, t_dev, ctx_dev, SEQ_LEN, packed=False)

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        cd = ctx_host.to(dev, non_blocking=True)
        td = t_host.to(dev, non_blocking=True)
        out = model([xd], td, [cd], seq_len=SEQ_LEN, flag=False)[0]
        out_host.copy_(out, non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    n_warm = args.warmup if args.quick else max(3, args.warmup)
    for _ in range(n_warm):
        step_device()
    torch.cuda.synchronize()

    with ClockSampler(local_rank) as clocks:
        ops.reset_launch_count()
        eng.timer.reset()
        eng.timer.active = True
        total_ms = timed(step_device, args.steps)
        eng.timer.active = False
        launches = ops.launch_count()
        kern = eng.timer.summary()
    if args.quick:
        e2e_ms = float("nan")
    else:
        for _ in range(2):
            step_e2e()
        e2e_ms = timed(step_e2e, args.steps)

    ms_per_step = total_ms / args.steps
    e2e_ms_per_step = e2e_ms / args.steps
    frames = LATENT[1]
    peaks = measured_peaks()
    C, F, L = cfg["dim"], cfg["ffn_dim"], SEQ_LEN
    step_flops = cfg["num_layers"] * block_flops(L, C, F, CTX_LEN)
    # dominant kernel: self-attention (4*L^2*C algorithmic FLOP per launch; with P GPUs each rank does 1/P of the heads)
    att = kern.get("self_attention")
    att_flops = 4.0 * L * L * C / world
    roof = None
    if att:
        ach = att_flops / (att["mean_ms"] * 1e-3) / 1e12
        peak = peaks["bf16_sustained"] or peaks["bf16_tflops"]
        roof = {"bound": "tensor", "kernel": "attention_kernel<true> (self-attention)", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one N=1 launch from the ncu --set full capture
                # (profiles/r01_final_ncu_attention.md: 343.6 MB + 104.6 MB; algorithmic Q+K+V+O bytes = 454 MB)
                "traffic": (448.2e6 if world == 1 else None), "traffic_unit": "bytes/launch",
                "peak_source": peaks["source"] + ", sustained cuBLAS bf16 (kernel timed inside a long step)",
                "mean_launch_ms": att["mean_ms"], "launches_timed": att["count"],
                "share_of_step": att["total_ms"] / total_ms}
    gemm_flops = {"gemm_qkv": 6.0 * L * C * C, "gemm_o": 2.0 * L * C * C, "gemm_ffn1": 2.0 * L * C * F,
                  "gemm_ffn2": 2.0 * L * C * F}
    kernels = {}
    for tag, s in kern.items():
        k = {"mean_ms": s["mean_ms"], "count": s["count"], "share_of_step": s["total_ms"] / total_ms}
        if tag in gemm_flops:
            k["tflops"] = gemm_flops[tag] / world / (s["mean_ms"] * 1e-3) / 1e12
        if tag == "ln_modulate":
            k["gbps"] = (L / world) * C * 6 / (s["mean_ms"] * 1e-3) / 1e9
        if tag == "rmsnorm_rope":
            k["gbps"] = (L / world) * C * 4 / (s["mean_ms"] * 1e-3) / 1e9
        kernels[tag] = k

    line = None
    if rank == 0:
        line = {
            "metric": "latent_frames_per_sec", "value": frames / (ms_per_step * 1e-3), "unit": "latent-frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "Yume-5B-720P single denoise step (WanModel.forward, flag=False), 81-frame 704x1280 "
                                   "latent [48,21,44,80], L=18480 tokens, 512-token text context, t=500",
                       "model": "Yume-5B-720P (Wan2.2-TI2V-5B geometry: dim 3072, ffn 14336, 24 heads, 30 layers), random init",
                       "parallelism": "single GPU" if world == 1 else
                       f"ulysses sp{world} ({'NVLink peer-memory exchange fused into kernels' if eng._sp_p2p else 'NCCL all-to-all'})",
                       "l2": "per-step working set (10 GB of bf16 weights + 1.5 GB activations) >> 126 MB L2; no flush needed",
                       "step_tflop": step_flops / 1e12},
            "step_tflops_achieved": step_flops / (ms_per_step * 1e-3) / 1e12,
            "e2e": {"value": frames / (e2e_ms_per_step * 1e-3), "unit": "latent-frames/s", "ms_per_step": e2e_ms_per_step,
                    "h2d_bytes_per_step": x_host.numel() * 4 + ctx_host.numel() * 2 + 4,
                    "d2h_bytes_per_step": out_host.numel() * 4,
                    "api": "yume_b200.model.WanModel5B.forward(x, t, context, seq_len, flag=False) with pinned host tensors"},
            "gpu_launches": launches,
            "roofline": roof,
            "kernels": kernels,
            "clocks": clocks.summary(),
        }
    if world == 1 and not args.no_cpu_baseline and not args.quick and rank == 0:
        r = cpu_reference_sample(budget_s=25.0, reps=1, warmup=1)
        line["cpu_baseline"] = {"value": frames / r["t_step_s"], "unit": "latent-frames/s", "cores": r["cores"],
                                "kind": "port", "sample": r["sample"], "ms_per_step": r["t_step_s"] * 1e3}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def chunk_arm(args):
    """Supplementary workloads (BASELINE.json configs[2]/[3]): one FramePack-chunk denoise forward.
      5b-chunk : Yume-5B, 13 latent frames @44x80 (5 history + 8 new), L = 9460, per-token t (0 / 900)
      14b-chunk: Yume-I2V-540P (14B), 13 latent frames @68x120, latent_frame_zero 8, L = 21930, CLIP + text context"""
    from yume_b200 import ops
    from yume_b200.dit import WanDiT

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    five = args.config == "5b-chunk"
    cfg = CFG_5B if five else CFG_14B
    sd = synthetic_state_dict(cfg, dev, seed=0)
    kw = {k: cfg[k] for k in ("dim", "ffn_dim", "num_heads", "num_layers", "in_dim", "out_dim", "text_len", "freq_dim")}
    eng = WanDiT(sd, cfg["variant"], device=dev, **kw)
    del sd
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev).manual_seed(1)
    ctx = torch.randn(CTX_LEN, cfg["text_dim"], generator=g, device=dev).to(torch.bfloat16)
    if five:
        x = torch.randn(48, 13, 44, 80, generator=g, device=dev)
        fwd = lambda: eng.forward(x, torch.tensor([[0.0, 900.0]], device=dev), ctx, 0, latent_frame_zero=8, packed=True)  # noqa: E731
        L, S = 9460, CTX_LEN
    else:
        x = torch.randn(16, 13, 68, 120, generator=g, device=dev)
        y = torch.randn(20, 13, 68, 120, generator=g, device=dev)
        clip = torch.randn(1, 257, 1280, generator=g, device=dev)
        fwd = lambda: eng.forward(x, torch.tensor([500.0], device=dev), ctx, 0, y=y, clip_fea=clip, latent_frame_zero=8,  # noqa: E731
                                  packed=True)
        L, S = 21930, CTX_LEN + 257
    for _ in range(max(3, args.warmup)):
        out = fwd()
    torch.cuda.synchronize()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flops = cfg["num_layers"] * block_flops(L, cfg["dim"], cfg["ffn_dim"], S)
    print(json.dumps({"metric": "latent_frames_per_sec", "value": 8 / (ms * 1e-3), "unit": "latent-frames/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": f"{args.config}: FramePack chunk forward, L={L}, out {tuple(out.shape)}",
                                 "step_tflop": flops / 1e12},
                      "step_tflops_achieved": flops / (ms * 1e-3) / 1e12, "gpu_launches": ops.launch_count(),
                      "finite": bool(torch.isfinite(out).all())}), flush=True)


def vae_arm(args):
    """Supplementary workload (BASELINE.json configs[4]): hyvideo causal 3D VAE tiled decode of z [1,16,21,90,160]
    (81 frames 720x1280) with the upstream 884-16c config, random-init weights. Prints the same JSON schema."""
    import math as _m

    from yume_b200 import ops
    from yume_b200.vae import CONFIG_884_16C, HyVaeDecoder, decoder_param_shapes

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for name, shape in decoder_param_shapes().items():
        t = torch.randn(shape, generator=g, device=dev)
        if name.endswith(".bias"):
            t = 0.05 * t
        elif len(shape) == 1:
            t = 1 + 0.1 * t
        else:
            t = t * (1.5 / _m.sqrt(_m.prod(shape[1:])))
        sd[name] = t
    eng = HyVaeDecoder(sd, device=dev, **CONFIG_884_16C)
    eng.enable_tiling()
    T, H, W = (21, 90, 160) if not args.quick else (5, 40, 40)
    z = torch.randn(1, 16, T, H, W, generator=g, device=dev)
    for _ in range(args.warmup):
        eng.decode(z)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.decode(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    frames = out.shape[2]
    print(json.dumps({"metric": "decoded_frames_per_sec", "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": f"hyvideo AutoencoderKLCausal3D tiled decode z[1,16,{T},{H},{W}] -> {tuple(out.shape)}",
                                 "tiles": "spatial 32x32 stride 24, temporal 17 stride 12 (upstream 884-16c)"},
                      "gpu_launches": ops.launch_count(), "finite": bool(torch.isfinite(out).all())}), flush=True)


def vae22_arm(args):
    """Supplementary workloads, random-init weights at the real widths:
    vae22 — Wan2.2 VAE decode of the 5B sampler's latent z [48,21,44,80] -> video [3,81,704,1280] (`Wan2_2_VAE.decode`);
    vae21 — Wan2.1 VAE decode of the 14B sampler's 540p latent z [16,21,68,120] -> [3,81,544,960] (`WanVAE.decode`)."""
    from yume_b200 import ops
    if args.workload == "vae22":
        from yume_b200.vae22 import Wan22VaeDecoder as Decoder, decoder_param_shapes
        zc, full, label = 48, (21, 44, 80), "Wan2.2 VAE"
    else:
        from yume_b200.vae21 import Wan21VaeDecoder as Decoder, decoder_param_shapes
        zc, full, label = 16, (21, 68, 120), "Wan2.1 VAE"

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for name, shape in decoder_param_shapes().items():
        t = torch.randn(shape, generator=g, device=dev)
        if name.endswith(".bias"):
            t = 0.05 * t
        elif name.endswith(".gamma"):
            t = 1 + 0.1 * t
        else:
            t = t * (0.8 / math.sqrt(math.prod(shape[1:])))
        sd[name] = t
    eng = Decoder(sd, device=dev)
    T, H, W = full if not args.quick else (5, 16, 16)
    z = torch.randn(zc, T, H, W, generator=g, device=dev)
    for _ in range(args.warmup):
        eng.decode(z)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.decode(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "decoded_frames_per_sec", "value": out.shape[1] / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": f"{label} whole-sequence decode z[{zc},{T},{H},{W}] -> {tuple(out.shape)}"},
                      "gpu_launches": ops.launch_count(), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                      "finite": bool(torch.isfinite(out).all())}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="dit", choices=["dit", "vae", "vae22", "vae21"])
    ap.add_argument("--config", default="5b-720p", choices=["5b-720p", "5b-chunk", "14b-chunk"],
                    help="5b-720p is the headline workload (BASELINE.json configs[1]); the others are supplementary")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="yume_b200", choices=["yume_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work for the whole run")
    ap.add_argument("--sp-transport", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--quick", action="store_true", help="profiling runs: exact --warmup, no e2e leg, no CPU baseline")
    args = ap.parse_args()
    if args.workload == "vae":
        vae_arm(args)
    elif args.workload in ("vae22", "vae21"):
        vae22_arm(args)
    elif args.config != "5b-720p":
        chunk_arm(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        product_arm(args)


if __name__ == "__main__":
    main()
