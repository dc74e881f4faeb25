"""Debug run of the lookahead attention kernel against a build with YB_DEBUG_WAIT (prints the barrier that timed out)."""
import sys
from pathlib import Path
sys.path.insert(0, ".")
import torch
from yume_b200 import _lib
_lib._LIB_PATH = Path("tools/build/libyume_b200_dbg.so")
from yume_b200 import ops
dev = "cuda"
for (Lq, Lk, heads) in [(256, 64, 1), (256, 128, 1), (256, 256, 1), (300, 1000, 2)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    q = torch.randn(Lq, heads * 128, generator=g).to(dev).bfloat16()
    k = torch.randn(Lk, heads * 128, generator=g).to(dev).bfloat16()
    v = torch.randn(Lk, heads * 128, generator=g).to(dev).bfloat16()
    ref = ops.attention(q, k, v, torch.empty_like(q), heads, softmax=0)
    torch.cuda.synchronize()
    print("case", Lq, Lk, heads, "classic ok", flush=True)
    out = ops.attention(q, k, v, torch.zeros_like(q), heads, softmax=2)
    torch.cuda.synchronize()
    print("   lookahead rel", float((out.float() - ref.float()).norm() / ref.float().norm()), flush=True)
