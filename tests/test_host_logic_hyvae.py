"""Host logic of the hyvideo VAE decoder engine (yume_b200/vae.py) on CPU: weight re-packing, the tile plan (temporal windows, row /
column starts, crop and cross-fade lengths, frame offsets), the raw-tile decode loop — over the torch stand-in for the C-ABI calls
(tests/helpers/torch_ops.py, whose tile assembly is the reference's own blend_v -> blend_h -> crop -> cat -> blend_t sequence) against
the fixtures generated from the reference-shaped decoder. The assembly KERNEL itself is checked by the `-m gpu` twin of this test."""
import pytest
import torch

from helpers import torch_ops
from oracle import hyvae
from yume_b200 import vae


@pytest.fixture()
def cpu_ops(monkeypatch):
    monkeypatch.setattr(vae, "ops", torch_ops)


@pytest.mark.parametrize("case", ["untiled", "untiled_t1", "spatial_tiled", "temporal_spatial_tiled"])
def test_hyvideo_decoder_host_logic_reproduces_reference_fixture(cpu_ops, golden_dir, case):
    g = torch.load(golden_dir / "hyvae_tiny.pt", weights_only=False)
    sd = hyvae.make_state_dict(g["seed_w"], **g["cfg"])
    c = g["cases"][case]
    eng = vae.HyVaeDecoder(sd, sample_size=c["sample_size"], sample_tsize=c["sample_tsize"], device="cpu", **g["cfg"])
    eng.enable_tiling(c["tiling"])
    z = torch.randn(1, 16, c["T"], c["H"], c["W"], generator=torch.Generator().manual_seed(c["seed"]))
    out = eng.decode(z)
    assert tuple(out.shape) == c["shape"]
    for key, got in (("sample", out[..., ::3, ::3]), ("rowsum", out.sum(-1)), ("colsum", out.sum(-2))):
        assert float((got - c[key]).norm() / c[key].norm()) < 3e-2, key
