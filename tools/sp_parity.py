"""Ulysses parity on N GPUs (launch with torchrun): every golden forward case of both tiny models, sequence-parallel,
against the reference-generated outputs. Exit code != 0 on any mismatch."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import synth  # noqa: E402  (test infrastructure)
from yume_b200.dit import WanDiT  # noqa: E402

TOL = 5e-3   # same bar as the single-GPU golden tests (tests/test_gpu_parity.py MODEL_TOL)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if os.environ.get("YB_ATT_FORCE_SPLIT"):          # test hook: drive the KV split + peer-scatter combine path
        from yume_b200 import _lib
        _lib.load().yb_debug_force_split(int(os.environ["YB_ATT_FORCE_SPLIT"]))
    bad = ran = 0
    for fname in ("wan23_tiny.pt", "wan21_tiny.pt", "wan23_h8.pt", "wan21_h8.pt"):   # 2-head and 8-head models
        g = torch.load(ROOT / "tests" / "golden" / fname, weights_only=False)
        cfg = g["cfg"]
        if cfg["num_heads"] % dist.get_world_size():
            continue
        sd = synth.make_state_dict(cfg, g["seed_w"])
        kw = synth.oracle_kwargs(cfg)
        variant = kw.pop("variant")
        eng = WanDiT(sd, variant, device=dev, **kw)
        eng.enable_sequence_parallel(dist.group.WORLD, transport=os.environ.get("YB_SP_TRANSPORT", "auto"))
        for name, c in g["cases"].items():
            inp = synth.make_inputs(cfg, c["seed"], c["frames"], c["H"], c["W"], c["ctx_len"])
            if variant == "5b":
                out = eng.forward(inp["x"], torch.tensor(c["t"]), inp["context"], c["seq_len"], latent_frame_zero=c["lfz"],
                                  packed=c["flag"])
            else:
                out = eng.forward(inp["x"], torch.tensor(c["t"]), inp["context"], c["seq_len"], y=inp["y"],
                                  clip_fea=inp["clip_fea"], latent_frame_zero=c["lfz"], packed=c["rand_num_img"] >= 0.4)
            r = float((out.cpu() - c["out"]).norm() / c["out"].norm())
            ok = out.shape == c["out"].shape and bool(torch.isfinite(out).all()) and r < TOL
            bad += 0 if ok else 1
            ran += 1
            if rank == 0 or not ok:
                print(f"[rank {rank}] sp{dist.get_world_size()} ({eng.sp_transport}{'/p2p' if eng._sp_p2p else ''}) {name}: "
                      f"rel {r:.3e} {'ok' if ok else 'MISMATCH'}", flush=True)
    if ran == 0:
        bad += 1
        print(f"[rank {rank}] no golden model has heads divisible by world={dist.get_world_size()}", flush=True)
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
