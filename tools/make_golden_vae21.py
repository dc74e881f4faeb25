"""Generate tests/golden/wan21vae_tiny.pt by running the REFERENCE's chunked Wan2.1 VAE decode
(/root/reference/wan/modules/vae.py::WanVAE_.decode, frame-by-frame with the feature cache) on CPU at reduced width.
Authoring container only."""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import wan21vae  # noqa: E402

TINY = dict(dim=32, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_upsample=(True, True, False))


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    spec = importlib.util.spec_from_file_location("ref_vae2_1", "/root/reference/wan/modules/vae.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    seed = 778
    sd = wan21vae.make_state_dict(seed, **TINY)
    # temperal_downsample is the reverse of the decoder's temperal_upsample (vae.py:498)
    model = ref.WanVAE_(dim=TINY["dim"], z_dim=TINY["z_dim"], dim_mult=list(TINY["dim_mult"]), num_res_blocks=2, attn_scales=[],
                        temperal_downsample=list(TINY["temperal_upsample"][::-1]))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "conv1.")) for k in missing), (missing[:4], unexpected)
    model.eval()
    g = torch.Generator().manual_seed(6)
    mean, std = 0.3 * torch.randn(TINY["z_dim"], generator=g), 0.5 + torch.rand(TINY["z_dim"], generator=g)
    scale = [mean, 1.0 / std]
    gold = {"cfg": TINY, "seed_w": seed, "mean": mean, "std": std,
            "weight_abs_sum": float(sum(v.abs().sum() for v in sd.values())), "cases": {}}
    for i, (name, (T, H, W)) in enumerate([("t1", (1, 4, 8)), ("t2", (2, 4, 8)), ("t5", (5, 4, 8)), ("t3_wide", (3, 8, 12))]):
        z = torch.randn(TINY["z_dim"], T, H, W, generator=torch.Generator().manual_seed(500 + i))
        out = model.decode(z.unsqueeze(0), scale).float().clamp_(-1, 1).squeeze(0)       # WanVAE.decode :655-663
        gold["cases"][name] = dict(seed=500 + i, T=T, H=H, W=W, shape=tuple(out.shape), sample=out[..., ::3, ::3].clone(),
                                   rowsum=out.sum(-1), colsum=out.sum(-2))
        print(name, tuple(out.shape), float(out.abs().mean()), float((out.abs() >= 1).float().mean()))
    path = ROOT / "tests" / "golden" / "wan21vae_tiny.pt"
    torch.save(gold, path)
    print(path.name, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
