"""One GEMM shape, a few launches (for ncu): python scripts/one_gemm.py M N K [nt|nn|tn|tt] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image2text_b200 import ops  # noqa: E402

M, N, K = [int(x) for x in sys.argv[1:4]]
layout = sys.argv[4] if len(sys.argv) > 4 else "nt"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
ak, bk = layout[0] == "n", layout[1] == "t"
a = torch.randn((M, K) if ak else (K, M), device="cuda").bfloat16()
b = torch.randn((N, K) if bk else (K, N), device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(a, b, out=out, a_kmajor=ak, b_kmajor=bk)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"{M}x{N}x{K} {layout}: {min(ts):.1f} us (cold L2, best of {iters})")
