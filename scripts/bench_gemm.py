"""tcgen05 GEMM micro-benchmark at the training shapes: TFLOP/s vs MEASURED_PEAKS.json (bf16_tflops burst).

    python scripts/bench_gemm.py            # one JSON line per shape

Each shape is timed with CUDA events over back-to-back launches on operands that rotate through more than the L2.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2text_b200 import ops  # noqa: E402

SHAPES = [  # (M, N, K, a_kmajor, b_kmajor, label)
    (2048, 2304, 768, True, True, "c_attn fwd"),
    (2048, 3072, 768, True, True, "c_fc fwd"),
    (2048, 768, 3072, True, True, "mlp c_proj fwd"),
    (1984, 50257, 768, True, True, "lm_head fwd"),
    (1984, 768, 50264, True, False, "lm_head dgrad"),
    (768, 3072, 2048, False, False, "c_fc wgrad"),
    (8192, 3072, 768, True, True, "c_fc fwd B=32"),
    (16384, 3072, 768, True, True, "c_fc fwd B=64"),
    (8192, 8192, 8192, True, True, "8192^3"),
]


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak = peaks["bf16_tflops"]
    dev = "cuda"
    for M, N, K, ak, bk, label in SHAPES:
        nbuf = max(2, int(256e6 // ((M * K + N * K + M * N) * 2)) + 1)
        nbuf = min(nbuf, 16)
        As = [torch.randn((M, K) if ak else (K, M), device=dev).bfloat16() for _ in range(nbuf)]
        Bs = [torch.randn((N, K) if bk else (K, N), device=dev).bfloat16() for _ in range(nbuf)]
        Npad = (N + 7) // 8 * 8
        out = torch.empty((M, Npad), device=dev, dtype=torch.bfloat16)
        def run(i):
            ops.gemm(As[i % nbuf], Bs[i % nbuf], out=out, a_kmajor=ak, b_kmajor=bk, M=M, N=N, K=K, ldc=Npad)
        def timed(fn, iters=20):
            """GPU time per launch: the launches are captured into a CUDA graph so that the host (ctypes / dispatcher
            overhead of ~10-15 us per call) is out of the measurement."""
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

        us = timed(run)
        # cuBLAS for comparison (library baseline, not the product path)
        a2s = [a if ak else a.t() for a in As]
        b2s = [b.t() if bk else b for b in Bs]
        us_cublas = timed(lambda i: torch.matmul(a2s[i % nbuf], b2s[i % nbuf]))
        tf = 2.0 * M * N * K / us / 1e6
        print(json.dumps({"shape": label, "M": M, "N": N, "K": K, "us": round(us, 2), "tflops": round(tf, 1),
                          "frac_of_peak": round(tf / peak, 3), "cublas_us": round(us_cublas, 2),
                          "cublas_tflops": round(2.0 * M * N * K / us_cublas / 1e6, 1)}))


if __name__ == "__main__":
    main()
