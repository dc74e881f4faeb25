"""Tile-parallel hyvideo VAE decode on N GPUs (launch with torchrun): every tiled golden case decoded with the tiles spread over
the ranks, compared with the reference-generated fixture on every rank; then (optional, --full) the 720p 81-frame decode timed
against the single-GPU decode. Exit code != 0 on any mismatch."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import hyvae  # noqa: E402  (test infrastructure: seeded weights)
from yume_b200.vae import CONFIG_884_16C, HyVaeDecoder, decoder_param_shapes  # noqa: E402

TOL = 3e-2


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    g = torch.load(ROOT / "tests" / "golden" / "hyvae_tiny.pt", weights_only=False)
    sd = hyvae.make_state_dict(g["seed_w"], **g["cfg"])
    bad = 0
    for name, c in g["cases"].items():
        if not c["tiling"]:
            continue
        eng = HyVaeDecoder(sd, sample_size=c["sample_size"], sample_tsize=c["sample_tsize"], device=dev, **g["cfg"])
        eng.enable_tiling(True)
        eng.enable_tile_parallel(dist.group.WORLD)
        z = torch.randn(1, 16, c["T"], c["H"], c["W"], generator=torch.Generator().manual_seed(c["seed"]))
        out = eng.decode(z).cpu()
        r = max(float((got - c[key]).norm() / c[key].norm())
                for key, got in (("sample", out[..., ::3, ::3]), ("rowsum", out.sum(-1)), ("colsum", out.sum(-2))))
        ok = tuple(out.shape) == c["shape"] and r < TOL
        bad += 0 if ok else 1
        print(f"[rank {rank}] tile-parallel x{world} {name}: rel {r:.3e} {'ok' if ok else 'MISMATCH'}", flush=True)
    if "--full" in sys.argv:
        gen = torch.Generator(device=dev).manual_seed(0)
        sdf = {}
        for n_, shape in decoder_param_shapes().items():
            t = torch.randn(shape, generator=gen, device=dev)
            sdf[n_] = 0.05 * t if n_.endswith(".bias") else (1 + 0.1 * t if len(shape) == 1 else t * (1.5 / (torch.tensor(shape[1:]).prod().item() ** 0.5)))
        eng = HyVaeDecoder(sdf, device=dev, **CONFIG_884_16C)
        eng.enable_tiling(True)
        z = torch.randn(1, 16, 21, 90, 160, generator=gen, device=dev)
        dist.broadcast(z, 0)
        times = {}
        for mode in ("single", "parallel"):
            if mode == "parallel":
                eng.enable_tile_parallel(dist.group.WORLD)
            out = eng.decode(z)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.decode(z)
            e1.record()
            dist.barrier(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            times[mode] = float(ms)
            ref = out if mode == "single" else ref
            if mode == "parallel":
                # not bit-identical run to run: GroupNorm statistics are fp64 atomics (summation order), a last-bit change of a
                # scale can flip bf16 roundings downstream
                d = float((out - ref).norm() / ref.norm())
                bad += 0 if d < 2e-3 else 1
                if rank == 0:
                    print(f"full 720p decode: single GPU {times['single']:.1f} ms, tile-parallel x{world} {times['parallel']:.1f} ms "
                          f"({times['single'] / times['parallel']:.2f}x), rel-Frobenius(parallel, single) {d:.2e}", flush=True)
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
