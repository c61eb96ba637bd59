"""Does alternating a small-shared-memory kernel with the 225 KB-shared-memory GEMM cost extra (SM carve-out switch)?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from image2text_b200 import ops  # noqa: E402

M, N, K = 64, 768, 768
nbuf = 64
As = [torch.randn((M, K), device="cuda").bfloat16() for _ in range(8)]
Bs = [torch.randn((N, K), device="cuda").bfloat16() for _ in range(nbuf)]
bias = torch.randn(N, device="cuda")
out = torch.zeros((M, N), device="cuda")
x = torch.randn((M, K), device="cuda")
g1, b1 = torch.ones(K, device="cuda"), torch.zeros(K, device="cuda")
small = torch.zeros(64, device="cuda")


def timed(fn, iters=128):
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


def gemm(i):
    ops.gemm(As[i % 8], Bs[i % nbuf], bias=bias, residual=out, out=out)


def ln(i):
    ops.layernorm(x, g1, b1, 1e-5, out_dtype=torch.bfloat16)


def fill(i):
    small.add_(1.0)


print("gemm alone        %.2f us" % timed(gemm))
print("ln alone          %.2f us" % timed(ln))
print("fill alone        %.2f us" % timed(fill))
print("ln + gemm         %.2f us" % timed(lambda i: (ln(i), gemm(i))))
print("fill + gemm       %.2f us" % timed(lambda i: (fill(i), gemm(i))))
print("gemm + gemm       %.2f us" % timed(lambda i: (gemm(i), gemm(i + 1))))
