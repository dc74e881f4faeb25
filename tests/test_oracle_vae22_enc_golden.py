"""Pin oracle/wan22vae_enc.py (whole-sequence form) against the reference's own CHUNKED Wan2.2 VAE encode."""
import pytest
import torch

from oracle import wan22vae_enc

TOL = 3e-5


@pytest.fixture(scope="module")
def gold(golden_dir):
    g = torch.load(golden_dir / "wan22vae_enc_tiny.pt", weights_only=False)
    sd = wan22vae_enc.make_state_dict(g["seed_w"], **g["cfg"])
    got = float(sum(v.abs().sum() for v in sd.values()))
    if abs(got - g["weight_abs_sum"]) > 1e-3 * g["weight_abs_sum"]:
        pytest.skip("torch CPU RNG stream differs from the one that generated the golden weights")
    return g, sd


@pytest.mark.parametrize("case", ["t1", "t5", "t9", "t17_wide"])
def test_whole_sequence_encode_equals_chunked_reference(gold, case):
    g, sd = gold
    c = g["cases"][case]
    m = wan22vae_enc.Wan22VaeEncodeOracle(sd, mean=g["mean"], std=g["std"], **g["cfg"])
    x = torch.randn(3, c["T"], c["H"], c["W"], generator=torch.Generator().manual_seed(c["seed"])).clamp_(-1, 1)
    mu = m.encode(x)
    assert tuple(mu.shape) == c["shape"]
    assert float((mu - c["mu"]).norm() / c["mu"].norm()) < TOL


def test_patchify_and_avgdown_match_their_definitions():
    x = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).view(1, 2, 3, 4, 6)
    p = wan22vae_enc.patchify2(x)
    # 'b c f (h q) (w r) -> b (c r q) f h w': channel index = (c*2 + r)*2 + q
    for c in range(2):
        for r in range(2):
            for q in range(2):
                assert torch.equal(p[0, (c * 2 + r) * 2 + q], x[0, c, :, q::2, r::2])
    y = torch.randn(1, 4, 3, 4, 4)
    a = wan22vae_enc.avg_down(y, 8, 2, 2)                       # T = 3 -> one zero frame in front -> 2 output frames
    assert a.shape == (1, 8, 2, 2, 2)
    ypad = torch.cat([torch.zeros(1, 4, 1, 4, 4), y], 2)
    # output channel o averages group_size = 4*8/8 = 4 consecutive entries of the (c, ft, fs_h, fs_w)-major expansion
    exp = ypad.view(1, 4, 2, 2, 2, 2, 2, 2).permute(0, 1, 3, 5, 7, 2, 4, 6).reshape(1, 32, 2, 2, 2).view(1, 8, 4, 2, 2, 2).mean(2)
    assert torch.allclose(a, exp)
