"""Dump the decode-side state-dict keys/shapes of the reference's Wan2.2 and Wan2.1 VAE module trees at their REAL sizes
(built on the meta device, no weights) to tests/golden/wan_vae_shapes.json. Authoring container only."""
import importlib.util
import json
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    r22 = load("/root/reference/wan23/modules/vae2_2.py", "ref_vae2_2")
    r21 = load("/root/reference/wan/modules/vae.py", "ref_vae2_1")
    with torch.device("meta"):
        # constructor arguments of `_video_vae` as the wrappers call it (vae2_2.py:876-906, 1030-1040; vae.py:571-598)
        m22 = r22.WanVAE_(dim=160, dec_dim=256, z_dim=48, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                          temperal_downsample=[False, True, True])
        m21 = r21.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                          temperal_downsample=[False, True, True])
    out = {}
    for name, m in (("wan22", m22), ("wan21", m21)):
        out[name] = {k: list(v.shape) for k, v in m.state_dict().items() if k.startswith(("decoder.", "conv2."))}
    path = ROOT / "tests" / "golden" / "wan_vae_shapes.json"
    path.write_text(json.dumps(out, indent=0, sort_keys=True))
    print(path, path.stat().st_size, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
