"""Generate tests/golden/wan21vae_enc_tiny.pt by running the REFERENCE's chunked Wan2.1 VAE encode
(/root/reference/wan/modules/vae.py::WanVAE_.encode: frame 0, then 4 frames per call, feature cache) on CPU at reduced
width. Authoring container only."""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import wan21vae_enc  # noqa: E402

TINY = dict(dim=32, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True))


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    spec = importlib.util.spec_from_file_location("ref_vae2_1", "/root/reference/wan/modules/vae.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    seed = 780
    sd = wan21vae_enc.make_state_dict(seed, **TINY)
    model = ref.WanVAE_(dim=TINY["dim"], z_dim=TINY["z_dim"], dim_mult=list(TINY["dim_mult"]), num_res_blocks=2,
                        attn_scales=[], temperal_downsample=list(TINY["temperal_downsample"]))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("decoder.", "conv2.")) for k in missing), (missing[:4], unexpected)
    model.eval()
    g = torch.Generator().manual_seed(7)
    mean, std = 0.3 * torch.randn(TINY["z_dim"], generator=g), 0.5 + torch.rand(TINY["z_dim"], generator=g)
    scale = [mean, 1.0 / std]
    gold = {"cfg": TINY, "seed_w": seed, "mean": mean, "std": std,
            "weight_abs_sum": float(sum(v.abs().sum() for v in sd.values())), "cases": {}}
    for i, (name, (T, H, W)) in enumerate([("t1", (1, 16, 32)), ("t5", (5, 16, 32)), ("t9", (9, 16, 16)), ("t17_wide", (17, 16, 48))]):
        x = torch.randn(3, T, H, W, generator=torch.Generator().manual_seed(700 + i)).clamp_(-1, 1)
        out = model.encode(x.unsqueeze(0), scale).float().squeeze(0)                # Wan2_2_VAE.encode :645-653
        gold["cases"][name] = dict(seed=700 + i, T=T, H=H, W=W, shape=tuple(out.shape), mu=out.clone())
        print(name, tuple(out.shape), float(out.abs().mean()))
    path = ROOT / "tests" / "golden" / "wan21vae_enc_tiny.pt"
    torch.save(gold, path)
    print(path.name, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
