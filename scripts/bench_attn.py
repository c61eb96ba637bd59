"""Self-attention forward / backward kernel timings (bf16, head_dim 64) at the shapes of the training step: CUDA events around N
back-to-back calls, per attention mode (1: tcgen05 forward + backward, 2: mma.sync).  One JSON line per (shape, mode).

    python scripts/bench_attn.py [--batches 8,64] [--T 256] [--heads 12] [--dropout 0.1]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="8,64")
    ap.add_argument("--T", type=int, default=256)
    ap.add_argument("--heads", type=int, default=12)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--mask", type=int, default=1, help="0 none, 1 causal, 2 prompt")
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    from image2text_b200 import ops
    from image2text_b200._lib import lib
    H, T, hs = args.heads, args.T, 64
    C = H * hs
    state = torch.tensor([1234, 7], dtype=torch.int64, device="cuda")
    site = ops.DropSite(args.dropout, state, 3) if args.dropout > 0 else None
    for B in [int(b) for b in args.batches.split(",")]:
        g = torch.Generator(device="cuda").manual_seed(B)
        qkv = (torch.randn(B * T, 3 * C, device="cuda", generator=g) * 0.5).bfloat16()
        dout = torch.randn(B * T, C, device="cuda", generator=g).bfloat16()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for mode in (1, 2):
            lib().i2t_set_tensor_core_attention(mode)
            try:
                out, lse = ops.attention_packed(qkv, B, T, H, args.mask, 0, want_lse=True, drop=site)
                res = {}
                for name, fn in (("fwd", lambda: ops.attention_packed(qkv, B, T, H, args.mask, 0, want_lse=True, drop=site)),
                                 ("bwd", lambda: ops.attention_packed_bwd(qkv, out, dout, lse, B, T, H, args.mask, 0, drop=site))):
                    for _ in range(5):
                        fn()
                    tot = 0.0
                    for _ in range(args.iters):
                        flush.zero_()                                   # inputs out of L2 between timed calls
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        fn()
                        e1.record()
                        torch.cuda.synchronize()
                        tot += e0.elapsed_time(e1)
                    res[name + "_us"] = round(tot / args.iters * 1e3, 1)
                # full (unmasked) FLOP count as SURVEY 8(d): fwd 4 T^2 C per sequence, bwd 2.5x
                fl = 4.0 * T * T * C * B
                res["fwd_tflops"] = round(fl / res["fwd_us"] / 1e6, 1)
                res["bwd_tflops"] = round(2.5 * fl / res["bwd_us"] / 1e6, 1)
                print(json.dumps(dict(B=B, T=T, H=H, mask=args.mask, dropout=args.dropout, mode=mode, **res)))
            finally:
                lib().i2t_set_tensor_core_attention(1)


if __name__ == "__main__":
    main()
