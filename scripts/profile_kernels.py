"""Launch the tensor-core kernels a few times at training shapes (run under ncu: scripts/gpu_ncu_tc.sh)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import ops  # noqa: E402

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if which == "gemm":
    for M, N, K in ((16384, 3072, 768), (8192, 8192, 8192)):
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, b, out=out)
else:
    B, T, H, hs = 64, 256, 12, 64
    qkv = torch.randn(B * T, 3 * H * hs, device=dev).bfloat16()
    dout = torch.randn(B * T, H * hs, device=dev).bfloat16()
    for _ in range(3):
        out, lse = ops.attention_packed(qkv, B, T, H, ops.MASK_PROMPT, 8, want_lse=True)
        ops.attention_packed_bwd(qkv, out, dout, lse, B, T, H, ops.MASK_PROMPT, 8)
torch.cuda.synchronize()
