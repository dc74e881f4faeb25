"""bench.py contract pieces that run without a GPU: the algorithmic-FLOP table of SURVEY.md §8(d), and the `--impl reference` line (the
CPU arm: rank 0 only, fixed sample, the base contract's keys) on a shrunken sample so that it runs in seconds."""
import json
import sys
import types
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_algorithmic_flops_match_the_survey_table():
    # 5B full 81f 290.4 TF / forward; 5B FramePack chunk 117.5; 14B FramePack (lfz 8) 935; 14B full 81f 2558
    c5, c14 = bench.CFG_5B, bench.CFG_14B
    f5 = lambda L: c5["num_layers"] * bench.block_flops(L, c5["dim"], c5["ffn_dim"], 512) / 1e12  # noqa: E731
    f14 = lambda L: c14["num_layers"] * bench.block_flops(L, c14["dim"], c14["ffn_dim"], 512 + 257) / 1e12  # noqa: E731
    assert abs(f5(18480) - 290.4) < 0.2 and abs(f5(9460) - 117.5) < 0.2
    assert abs(f14(21930) - 935) < 2 and abs(f14(42840) - 2558) < 4
    assert bench.SEQ_LEN == 18480 and bench.LATENT == (48, 21, 44, 80)


def test_reference_arm_line_keeps_the_contract(monkeypatch, capsys):
    monkeypatch.setattr(bench, "CPU_SAMPLE_L", 110)
    monkeypatch.setattr(bench, "CPU_SAMPLE_GRID", (1, 11, 10))
    monkeypatch.setattr(bench, "CPU_SAMPLE_REPS", 2)
    monkeypatch.setenv("RANK", "0")
    args = types.SimpleNamespace(gpus=1, steps=3, warmup=1)
    bench.reference_arm(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["steps"] == 3 and line["warmup"] == 1 and line["n_gpus"] == 1 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and len(cb["sample_seconds"]) == 2
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and abs(line["ms_per_step"] - 1e3 * bench.LATENT[1] / line["value"]) < 1e-6 * line["ms_per_step"]


def test_reference_arm_is_silent_on_other_ranks(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "3")
    bench.reference_arm(types.SimpleNamespace(gpus=8, steps=1, warmup=0))
    assert capsys.readouterr().out == ""
