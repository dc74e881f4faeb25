"""Pin oracle/wan21vae_enc.py (whole-sequence form) against the reference's own CHUNKED Wan2.1 VAE encode."""
import pytest
import torch

from oracle import wan21vae_enc

TOL = 3e-5


@pytest.fixture(scope="module")
def gold(golden_dir):
    g = torch.load(golden_dir / "wan21vae_enc_tiny.pt", weights_only=False)
    sd = wan21vae_enc.make_state_dict(g["seed_w"], **g["cfg"])
    got = float(sum(v.abs().sum() for v in sd.values()))
    if abs(got - g["weight_abs_sum"]) > 1e-3 * g["weight_abs_sum"]:
        pytest.skip("torch CPU RNG stream differs from the one that generated the golden weights")
    return g, sd


@pytest.mark.parametrize("case", ["t1", "t5", "t9", "t17_wide"])
def test_whole_sequence_encode_equals_chunked_reference(gold, case):
    g, sd = gold
    c = g["cases"][case]
    m = wan21vae_enc.Wan21VaeEncodeOracle(sd, mean=g["mean"], std=g["std"], **g["cfg"])
    x = torch.randn(3, c["T"], c["H"], c["W"], generator=torch.Generator().manual_seed(c["seed"])).clamp_(-1, 1)
    mu = m.encode(x)
    assert tuple(mu.shape) == c["shape"]
    assert float((mu - c["mu"]).norm() / c["mu"].norm()) < TOL
