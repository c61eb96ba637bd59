"""Run the bf16 attention backward for one shape (argv: B T H mask) and compare with the mma.sync kernel; used under `timeout`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image2text_b200 import ops  # noqa: E402
from image2text_b200._lib import lib  # noqa: E402

B, T, H, mask = [int(x) for x in sys.argv[1:5]]
C = H * 64
g = torch.Generator(device="cuda").manual_seed(1)
qkv = (torch.randn(B * T, 3 * C, device="cuda", generator=g) * 0.5).bfloat16()
dout = torch.randn(B * T, C, device="cuda", generator=g).bfloat16()
out, lse = ops.attention_packed(qkv, B, T, H, mask, 0, want_lse=True)
torch.cuda.synchronize()
print("fwd ok", flush=True)
lib().i2t_set_tensor_core_attention(2)
ref = ops.attention_packed_bwd(qkv, out, dout, lse, B, T, H, mask, 0)
torch.cuda.synchronize()
lib().i2t_set_tensor_core_attention(1)
got = ops.attention_packed_bwd(qkv, out, dout, lse, B, T, H, mask, 0)
torch.cuda.synchronize()
for name, a, b in zip("qkv", got.float().split(C, 1), ref.float().split(C, 1)):
    print(name, float((a - b).norm() / b.norm()), flush=True)
if os.environ.get("I2T_ATTN_BWD_DEBUG") == "-1":
    import ctypes
    buf = (ctypes.c_longlong * 32)()
    from image2text_b200._lib import call
    call("i2t_attn_bwd_trace", ctypes.addressof(buf))
    t0 = buf[0]
    names = {0: "start"}
    for t in range(3):
        names.update({1 + 4 * t: f"mma  pair{t} operands ready", 2 + 4 * t: f"mma  pair{t} S/dP issued", 3 + 4 * t: f"mma  pair{t} P/dS arrived",
                      4 + 4 * t: f"mma  pair{t} dV/dK/dQ issued", 17 + 4 * t: f"soft pair{t} S/dP complete", 18 + 4 * t: f"soft pair{t} P/dS written",
                      19 + 4 * t: f"soft pair{t} mma2 complete", 20 + 4 * t: f"soft pair{t} dK/dV stored"})
    names.update({16: "soft row scalars read", 31: "soft dQ stored", 13: "barriers + TMEM ready", 14: "predecessor complete",
                  15: "tma  all loads issued", 29: "soft dK/dV of key block 0 in registers", 30: "soft dK/dV of key block 1 in registers"})
    for i, v in sorted(((i, buf[i]) for i in range(32) if buf[i]), key=lambda kv: kv[1]):
        print(f"{(v - t0) / 1.9e3:8.2f} us  {names.get(i, i)}")
