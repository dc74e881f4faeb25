"""Self-attention comparators on one B200, same tensors, same shape (Yume-5B-720P: 24 heads x 128, L = 18 480, non-causal):
this repo's tcgen05 kernel, torch SDPA with the cuDNN backend, torch SDPA with the flash backend, and the flash_attn package (the
kernel family the reference calls, wan23/modules/attention.py:93-121). Prints ONE JSON object (-> profiles/r02_attention_comparators.json):
per entry the best-of-3 mean over 10 launches (CUDA events), TFLOP/s, SM clock under load and, for the fused kernels, the error against
fp32 SDPA on 256 sampled rows."""
import json
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from yume_b200 import ops  # noqa: E402
from yume_b200.utils import ClockSampler  # noqa: E402

dev = torch.device("cuda", 0)
heads, L = 24, 18480
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(L, 3 * heads * 128, generator=g, device=dev).bfloat16()
q, k, v = qkv[:, :heads * 128], qkv[:, heads * 128:2 * heads * 128], qkv[:, 2 * heads * 128:]
qh, kh, vh = (t.view(L, heads, 128).transpose(0, 1)[None].contiguous() for t in (q, k, v))     # [1, H, L, D]
qf, kf, vf = (t.view(1, L, heads, 128).contiguous() for t in (q, k, v))                          # [1, L, H, D]
out = torch.empty(L, heads * 128, device=dev, dtype=torch.bfloat16)
flop = 4.0 * L * L * heads * 128
idx = torch.randint(0, L, (256,), generator=g, device=dev)
ref = torch.nn.functional.scaled_dot_product_attention(qh[:, :, idx].float(), kh.float(), vh.float())[0].transpose(0, 1).reshape(256, -1)


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def entry(name, fn, result_rows):
    try:
        with ClockSampler(0) as clocks:
            ms = min(timeit(fn) for _ in range(3))
        got = result_rows()
        err = float((got.float() - ref).norm() / ref.norm())
        return {"name": name, "ms": ms, "tflops": flop / ms / 1e9, "rel_fro_vs_fp32_sdpa_256_rows": err, "clocks": clocks.summary()}
    except Exception as e:  # a backend missing on this box is a result, not a failure
        return {"name": name, "unavailable": f"{type(e).__name__}: {e}"[:300]}


res = []
res.append(entry("yume_b200 attention_kernel (tcgen05, P in TMEM)", lambda: ops.attention(q, k, v, out, heads), lambda: out[idx]))
from torch.nn.attention import SDPBackend, sdpa_kernel  # noqa: E402

holder = {}
for label, backend in (("torch SDPA, cuDNN backend", SDPBackend.CUDNN_ATTENTION), ("torch SDPA, flash backend", SDPBackend.FLASH_ATTENTION)):
    def run(backend=backend):
        with sdpa_kernel(backend):
            holder["o"] = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
    res.append(entry(label, run, lambda: holder["o"][0, :, idx].transpose(0, 1).reshape(256, -1)))


def run_default():
    holder["o"] = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)


res.append(entry("torch SDPA, default dispatch", run_default, lambda: holder["o"][0, :, idx].transpose(0, 1).reshape(256, -1)))
try:
    import flash_attn
    from flash_attn import flash_attn_func

    def run_fa():
        holder["o"] = flash_attn_func(qf, kf, vf)
    res.append(entry(f"flash_attn {flash_attn.__version__} flash_attn_func (the reference's kernel family)", run_fa,
                     lambda: holder["o"][0, idx].reshape(256, -1)))
except Exception as e:
    res.append({"name": "flash_attn package", "unavailable": f"{type(e).__name__}: {e}"[:300]})
print(json.dumps({"shape": {"heads": heads, "head_dim": 128, "Lq": L, "Lk": L, "causal": False, "dtype": "bf16"},
                  "flop_per_call": flop, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
                  "timing": "CUDA events, best of 3 x mean of 10 back-to-back launches, standalone (not inside a denoise step)",
                  "entries": res}, indent=1))
