"""Run-to-run reproducibility of the hyvideo tiled decode on one GPU: the full 720p / 81-frame latent decoded three times with the
same weights; prints rel-Frobenius and max |diff| between runs (0 / 0 = bit-reproducible). Exit code 1 above 1e-6."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from yume_b200.vae import CONFIG_884_16C, HyVaeDecoder, decoder_param_shapes  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
sd = {}
for n_, shape in decoder_param_shapes().items():
    t = torch.randn(shape, generator=gen, device=dev)
    sd[n_] = 0.05 * t if n_.endswith(".bias") else (1 + 0.1 * t if len(shape) == 1 else t * (1.5 / (torch.tensor(shape[1:]).prod().item() ** 0.5)))
eng = HyVaeDecoder(sd, device=dev, **CONFIG_884_16C)
eng.enable_tiling(True)
z = torch.randn(1, 16, 21, 90, 160, generator=gen, device=dev)
ref = eng.decode(z)
worst = 0.0
for i in range(2):
    out = eng.decode(z)
    d = float((out - ref).norm() / ref.norm())
    worst = max(worst, d)
    print(f"run {i + 1} vs run 0: rel-Frobenius {d:.3e}, max |diff| {float((out - ref).abs().max()):.3e}", flush=True)
sys.exit(1 if worst > 1e-6 else 0)
