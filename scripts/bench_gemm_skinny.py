"""Skinny (decode over a batch of sequences) and few-tile (weight-gradient) GEMM shapes: fp32 output, split-K on / off,
cuBLAS (bf16 out) beside it.  GPU time per launch from a CUDA graph of back-to-back launches over rotating operands."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import ops  # noqa: E402
from image2text_b200._lib import lib  # noqa: E402

SHAPES = [(64, 768, 768, True, True, "proj B=64"), (64, 2304, 768, True, True, "qkv B=64"), (64, 768, 3072, True, True, "mlp proj B=64"),
          (64, 3072, 768, True, True, "fc B=64"), (512, 768, 768, True, True, "proj B=512"), (512, 2304, 768, True, True, "qkv B=512"),
          (512, 768, 3072, True, True, "mlp proj B=512"), (768, 768, 16384, False, False, "c_proj wgrad B=64"),
          (2304, 768, 16384, False, False, "c_attn wgrad B=64"), (3072, 768, 16384, False, False, "c_fc wgrad B=64")]


def timed(fn, iters=160):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for M, N, K, ak, bk, label in SHAPES:
    # weights rotate through more than the 126 MB L2 for the decode shapes (every decode step reads each weight once, cold)
    nbuf = 8 if M > 512 else max(8, int(200e6 // (N * K * 2)))
    As = [torch.randn((M, K) if ak else (K, M), device="cuda").bfloat16() for _ in range(8)]
    Bs = [torch.randn((N, K) if bk else (K, N), device="cuda").bfloat16() for _ in range(nbuf)]
    out = torch.zeros((M, N), device="cuda")
    res = {}
    for split in (1, 0):
        lib().i2t_set_gemm_split_k(split)
        res[split] = timed(lambda i: ops.gemm(As[i % 8], Bs[i % nbuf], out=out, a_kmajor=ak, b_kmajor=bk, accumulate=True))
    lib().i2t_set_gemm_split_k(1)
    a2s = [a if ak else a.t() for a in As] * (nbuf // 8 + 1)
    b2s = [b.t() if bk else b for b in Bs]
    cub = timed(lambda i: torch.matmul(a2s[i % nbuf], b2s[i % nbuf]))
    wbytes = (N * K + M * K) * 2
    print(json.dumps({"shape": label, "M": M, "N": N, "K": K, "us_split": round(res[1], 2), "us_nosplit": round(res[0], 2),
                      "us_cublas_bf16out": round(cub, 2), "operand_GBps_split": round(wbytes / res[1] / 1e3, 1),
                      "tflops_split": round(2.0 * M * N * K / res[1] / 1e6, 1)}), flush=True)
