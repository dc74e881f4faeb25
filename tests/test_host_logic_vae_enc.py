"""Host logic of the Wan VAE ENCODER engines (yume_b200/vae_enc.py) on CPU: the engines' own code — weight re-packing, folded
latent normalisation, frame bookkeeping, strided Resample wiring, AvgDown3D shortcut — runs over a torch stand-in for the C-ABI
calls (tests/helpers/torch_ops.py, same layouts and argument meaning) and must reproduce the fixtures the reference's own chunked
encode generated. The CUDA kernels themselves are checked by the `-m gpu` twins of these tests."""
import pytest
import torch

from oracle import wan21vae_enc, wan22vae_enc
from helpers import torch_ops
from yume_b200 import vae22, vae_enc


@pytest.fixture()
def cpu_ops(monkeypatch):
    monkeypatch.setattr(vae22, "ops", torch_ops)
    monkeypatch.setattr(vae_enc, "ops", torch_ops)


def _gold(golden_dir, name, mod):
    g = torch.load(golden_dir / name, weights_only=False)
    sd = mod.make_state_dict(g["seed_w"], **g["cfg"])
    got = float(sum(v.abs().sum() for v in sd.values()))
    if abs(got - g["weight_abs_sum"]) > 1e-3 * g["weight_abs_sum"]:
        pytest.skip("torch CPU RNG stream differs from the one that generated the golden weights")
    return g, sd


@pytest.mark.parametrize("which,case", [(w, c) for w in ("wan21", "wan22") for c in ("t1", "t5", "t9", "t17_wide")])
def test_encoder_host_logic_reproduces_reference_fixture(cpu_ops, golden_dir, which, case):
    mod, Engine, name = ((wan21vae_enc, vae_enc.Wan21VaeEncoder, "wan21vae_enc_tiny.pt") if which == "wan21" else
                         (wan22vae_enc, vae_enc.Wan22VaeEncoder, "wan22vae_enc_tiny.pt"))
    g, sd = _gold(golden_dir, name, mod)
    c = g["cases"][case]
    eng = Engine(sd, mean=g["mean"], std=g["std"], device="cpu", **g["cfg"])
    x = torch.randn(3, c["T"], c["H"], c["W"], generator=torch.Generator().manual_seed(c["seed"])).clamp_(-1, 1)
    mu = eng.encode(x)
    assert tuple(mu.shape) == c["shape"]
    assert float((mu - c["mu"]).norm() / c["mu"].norm()) < 3e-2      # bf16 activations / weights against the fp32 reference


def test_encoder_truncates_to_1_plus_4k_frames(cpu_ops, golden_dir):
    g, sd = _gold(golden_dir, "wan21vae_enc_tiny.pt", wan21vae_enc)
    eng = vae_enc.Wan21VaeEncoder(sd, mean=g["mean"], std=g["std"], device="cpu", **g["cfg"])
    x = torch.randn(3, 8, 16, 16, generator=torch.Generator().manual_seed(1)).clamp_(-1, 1)
    assert torch.equal(eng.encode(x), eng.encode(x[:, :5]))          # `iter_ = 1 + (t - 1) // 4` (vae.py:520-521)
