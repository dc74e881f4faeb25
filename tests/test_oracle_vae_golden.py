"""Pin oracle/hyvae.py against outputs of the reference's own hyvideo/vae code (tools/make_golden_vae.py)."""
import pytest
import torch

from oracle import hyvae

TOL = 3e-5


@pytest.fixture(scope="module")
def gold(golden_dir):
    g = torch.load(golden_dir / "hyvae_tiny.pt", weights_only=False)
    sd = hyvae.make_state_dict(g["seed_w"], **g["cfg"])
    got = float(sum(v.abs().sum() for v in sd.values()))
    if abs(got - g["weight_abs_sum"]) > 1e-3 * g["weight_abs_sum"]:
        pytest.skip("torch CPU RNG stream differs from the one that generated the golden weights")
    return g, sd


def check_against_fixture(out, c, tol):
    assert tuple(out.shape) == c["shape"]
    for name, got in (("sample", out[..., ::3, ::3]), ("rowsum", out.sum(-1)), ("colsum", out.sum(-2))):
        want = c[name]
        rel = float((got.float().cpu() - want).norm() / want.norm())
        assert rel < tol, (name, rel)


@pytest.mark.parametrize("case", ["untiled", "untiled_t1", "spatial_tiled", "temporal_spatial_tiled"])
def test_vae_decode_matches_reference(gold, case):
    g, sd = gold
    c = g["cases"][case]
    m = hyvae.HyVaeOracle(sd, sample_size=c["sample_size"], sample_tsize=c["sample_tsize"], **g["cfg"])
    m.enable_tiling(c["tiling"])
    z = torch.randn(1, 16, c["T"], c["H"], c["W"], generator=torch.Generator().manual_seed(c["seed"]))
    check_against_fixture(m.decode(z), c, TOL)


def test_real_config_tiling_parameters():
    m = hyvae.HyVaeOracle({}, **hyvae.CONFIG_884_16C)
    assert (m.tile_latent_min_size, m.tile_sample_min_size, m.tile_latent_min_tsize, m.tile_sample_min_tsize) == (32, 256, 16, 64)
    assert m.up_factors == [(1, 2, 2), (2, 2, 2), (2, 2, 2), None]
