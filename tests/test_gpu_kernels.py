"""Kernel-level parity: every libi2t entry point against a plain fp32 (or fp64) PyTorch/oracle evaluation of the same
op on the same seeded inputs.  All tests call through the C ABI (ctypes) on cuda:0."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from image2text_b200 import ops  # noqa: E402
from image2text_b200._lib import call, lib  # noqa: E402
from image2text_b200.ops import ptr, stream  # noqa: E402
from oracle import i2t_oracle as O  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module", autouse=True)
def _exact_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    lib()
    yield


# ------------------------------------------------------------------ layernorm ---------------------------------
@pytest.mark.parametrize("rows,cols,eps", [(37, 768, 1e-5), (2048, 768, 1e-6), (5, 128, 1e-5), (9, 1600, 1e-5)])
def test_layernorm_fwd_bwd(rows, cols, eps):
    x = rnd(rows, cols, seed=1, scale=2.0) + 0.3
    g, b = 1 + 0.1 * rnd(cols, seed=2), 0.1 * rnd(cols, seed=3)
    dy = rnd(rows, cols, seed=4)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr.double(), (cols,), gr.double(), br.double(), eps)
    ref.backward(dy.double())
    y, mean, rstd = ops.layernorm(x.to(DEV), g.to(DEV), b.to(DEV), eps, want_stats=True)
    assert relerr(y, ref) < 2e-6
    dg = torch.zeros(cols, device=DEV)
    db = torch.zeros(cols, device=DEV)
    dx = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g.to(DEV), mean, rstd, dg, db)
    assert relerr(dx, xr.grad) < 5e-6
    assert relerr(dg, gr.grad) < 5e-6 and relerr(db, br.grad) < 5e-6
    # bf16 output, no bias, strided rows (token 0 of every image)
    y16 = ops.layernorm(x.to(DEV), g.to(DEV), None, eps, out_dtype=torch.bfloat16)
    ref16 = F.layer_norm(x, (cols,), g, None, eps)
    assert relerr(y16.float(), ref16) < 8e-3
    if rows % 4 == 1:
        ys = ops.layernorm(x.to(DEV), g.to(DEV), b.to(DEV), eps, rows=(rows + 3) // 4, row_stride=4 * cols)
        assert relerr(ys, ref[::4]) < 2e-6


# ------------------------------------------------------------------ GEMM --------------------------------------
@pytest.mark.parametrize("M,N,K", [(1576, 768, 768), (64, 2304, 768), (197, 517, 3072), (5, 50257, 128), (130, 131, 20)])
@pytest.mark.parametrize("layout", ["nt", "nn", "tn", "tt"])
def test_gemm_fp32_layouts(M, N, K, layout):
    if layout != "nt" and M * N * K > 2e9:
        pytest.skip("large case only for the forward layout")
    a, b = rnd(M, K, seed=5), rnd(N, K, seed=6, scale=0.05)
    bias, res = rnd(N, seed=7), rnd(M, N, seed=8)
    ref = (a.double() @ b.double().t())
    A = a.to(DEV) if layout[0] == "n" else a.t().contiguous().to(DEV)
    Bm = b.to(DEV) if layout[1] == "t" else b.t().contiguous().to(DEV)
    out = ops.gemm(A, Bm, a_kmajor=layout[0] == "n", b_kmajor=layout[1] == "t")
    assert relerr(out, ref) < 5e-6          # fp32 accumulation over K <= 3072 against an fp64 reference
    if layout == "nt":
        out = ops.gemm(A, Bm, bias=bias.to(DEV), residual=res.to(DEV), act=ops.ACT_GELU_TANH)
        ref2 = O.gelu_tanh(ref + bias.double()) + res.double()
        assert relerr(out, ref2) < 6e-6
        out = ops.gemm(A, Bm, bias=bias.to(DEV), act=ops.ACT_GELU_ERF)
        assert relerr(out, O.gelu_erf(ref + bias.double())) < 6e-6
        acc = res.to(DEV).clone()
        ops.gemm(A, Bm, out=acc, accumulate=True)
        assert relerr(acc, ref + res.double()) < 6e-6
        # in-place residual (C aliases residual), as the residual stream is updated
        x = res.to(DEV).clone()
        ops.gemm(A, Bm, bias=bias.to(DEV), residual=x, out=x)
        assert relerr(x, ref + bias.double() + res.double()) < 6e-6


@pytest.mark.parametrize("M,N,K", [(1576, 2304, 768), (128, 128, 64), (300, 50257, 768), (2048, 768, 3072), (1, 256, 768),
                                   (77, 200, 136)])
def test_gemm_bf16_tcgen05_and_fallback(M, N, K):
    a = rnd(M, K, seed=9).to(torch.bfloat16)
    b = rnd(N, K, seed=10, scale=0.05).to(torch.bfloat16)
    bias, res = rnd(N, seed=11), rnd(M, N, seed=12)
    ref = a.double() @ b.double().t()
    for tc in (1, 0):
        lib().i2t_set_tensor_core_gemm(tc)
        out = ops.gemm(a.to(DEV), b.to(DEV))
        assert relerr(out, ref) < 1e-5, f"tc={tc}"
        out = ops.gemm(a.to(DEV), b.to(DEV), bias=bias.to(DEV), residual=res.to(DEV), act=ops.ACT_GELU_TANH,
                       out_dtype=torch.float32)
        assert relerr(out, O.gelu_tanh(ref + bias.double()) + res.double()) < 1e-5, f"tc={tc}"
        out = ops.gemm(a.to(DEV), b.to(DEV), bias=bias.to(DEV), out_dtype=torch.bfloat16)
        assert relerr(out.float(), ref + bias.double()) < 8e-3, f"tc={tc}"
    lib().i2t_set_tensor_core_gemm(1)


@pytest.mark.parametrize("M,N,K", [(2048, 768, 3072), (1984, 768, 768), (768, 3072, 2048), (200, 136, 264), (128, 64, 72)])
@pytest.mark.parametrize("layout", ["nn", "tn", "tt"])
def test_gemm_bf16_tcgen05_mn_major_layouts(M, N, K, layout):
    """dgrad (A K-major, B MN-major), wgrad (both MN-major) and the A-transposed case on the tensor cores:
    MN-major UMMA descriptors + TMA boxes over the K-row-major storage."""
    a = rnd(M, K, seed=50).to(torch.bfloat16)
    b = rnd(N, K, seed=51, scale=0.05).to(torch.bfloat16)
    ref = a.double() @ b.double().t()
    A = (a if layout[0] == "n" else a.t().contiguous()).to(DEV)
    Bm = (b if layout[1] == "t" else b.t().contiguous()).to(DEV)
    for tc in (1, 0):
        lib().i2t_set_tensor_core_gemm(tc)
        out = ops.gemm(A, Bm, a_kmajor=layout[0] == "n", b_kmajor=layout[1] == "t")
        assert relerr(out, ref) < 1e-5, f"tc={tc}"
    lib().i2t_set_tensor_core_gemm(1)
    acc = torch.ones(M, N, device=DEV)
    ops.gemm(A, Bm, out=acc, a_kmajor=layout[0] == "n", b_kmajor=layout[1] == "t", accumulate=True)
    assert relerr(acc, ref + 1.0) < 1e-5
    # bf16 output (the dgrad that feeds the next backward kernel): N = 768 at M ~ 2048 takes the 128 x 96 tile whose bf16 boxes
    # are 64 bytes wide
    o16 = ops.gemm(A, Bm, a_kmajor=layout[0] == "n", b_kmajor=layout[1] == "t", out_dtype=torch.bfloat16)
    assert relerr(o16.float(), ref) < 8e-3


def test_gemm_bf16_96_column_tile_matches_the_128_column_tile():
    """The tile-shape switch (I2T_GEMM_TILE96) must not change results beyond fp32 summation order: same shapes through both tilings,
    every epilogue (bias, GELU, fp32 residual, bf16 / fp32 output, odd N tail)."""
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, ".")
from image2text_b200 import ops
g = torch.Generator(device="cuda").manual_seed(5)
outs = []
for (M, N, K) in [(2048, 768, 768), (1576, 768, 3072), (256, 200, 136), (300, 97, 64)]:
    a = (torch.randn(M, K, device="cuda", generator=g)).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    outs.append(ops.gemm(a, b, bias=bias, residual=res, act=ops.ACT_GELU_TANH, out_dtype=torch.float32))
    Np = (N + 7) // 8 * 8
    o = torch.zeros(M, Np, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, bias=bias, out=o[:, :N], ldc=Np)
    outs.append(o.float())
    bt = b.t().contiguous()
    outs.append(ops.gemm(a, bt, b_kmajor=False, out_dtype=torch.bfloat16).float() if N % 8 == 0 else o.float())
torch.save([o.cpu() for o in outs], sys.argv[1])
'''
    import os
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for flag in ("1", "0"):
            path = os.path.join(td, f"o{flag}.pt")
            subprocess.run([sys.executable, "-c", code, path], check=True, env=dict(os.environ, I2T_GEMM_TILE96=flag),
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            res[flag] = torch.load(path)
    for x, y in zip(res["1"], res["0"]):
        assert relerr(x, y) < 2e-3 and float((x - y).abs().max()) < 0.1


@pytest.mark.parametrize("M,N,K", [(4096, 2304, 768), (2000, 4000, 200), (768, 3072, 2048), (16384, 776, 136)])
@pytest.mark.parametrize("layout", ["nt", "nn", "tn", "tt"])
def test_gemm_bf16_cta_pair(M, N, K, layout):
    """CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 and 256 x 128 tiles) in all four operand layouts, with M / N / K
    tails, against fp64 and against the single-CTA kernel (A/B switch)."""
    a = rnd(M, K, seed=60).to(torch.bfloat16)
    b = rnd(N, K, seed=61, scale=0.05).to(torch.bfloat16)
    bias, res = rnd(N, seed=62), rnd(M, N, seed=63)
    ref = a.double() @ b.double().t()
    A = (a if layout[0] == "n" else a.t().contiguous()).to(DEV)
    Bm = (b if layout[1] == "t" else b.t().contiguous()).to(DEV)
    kw = dict(a_kmajor=layout[0] == "n", b_kmajor=layout[1] == "t")
    outs = {}
    for pair in (1, 0):
        lib().i2t_set_gemm_cta_pair(pair)
        try:
            out = ops.gemm(A, Bm, **kw)
            assert relerr(out, ref) < 1e-5, f"pair={pair}"
            outs[pair] = out
            out2 = ops.gemm(A, Bm, bias=bias.to(DEV), residual=res.to(DEV), act=ops.ACT_GELU_TANH, **kw)
            assert relerr(out2, O.gelu_tanh(ref + bias.double()) + res.double()) < 1e-5, f"pair={pair}"
        finally:
            lib().i2t_set_gemm_cta_pair(1)
    assert relerr(outs[1], outs[0].double()) < 2e-6       # same products, fp32 accumulation in a different tile order
    o16 = ops.gemm(A, Bm, bias=bias.to(DEV), out_dtype=torch.bfloat16, **kw)
    assert relerr(o16.float(), ref + bias.double()) < 8e-3


@pytest.mark.parametrize("M,N,K,layout", [(64, 768, 3072, "nt"), (64, 2304, 768, "nt"), (512, 768, 768, "nt"), (40, 776, 1000, "nt"),
                                            (768, 768, 8192, "tn"), (2304, 768, 4096, "tn"), (64, 768, 3072, "nn")])
def test_gemm_bf16_split_k(M, N, K, layout):
    """Few output tiles + long K (decode projections over a batch of sequences, weight gradients): K slices on otherwise
    idle SMs, partial tiles added with fp32 atomics.  Plain (zeroed C + bias on slice 0), in-place residual and accumulate
    forms, against fp64 and against the unsplit kernel (A/B switch)."""
    a = rnd(M, K, seed=70).to(torch.bfloat16)
    b = rnd(N, K, seed=71, scale=0.05).to(torch.bfloat16)
    bias, res = rnd(N, seed=72), rnd(M, N, seed=73)
    ref = a.double() @ b.double().t()
    A = (a if layout[0] == "n" else a.t().contiguous()).to(DEV)
    Bm = (b if layout[1] == "t" else b.t().contiguous()).to(DEV)
    kw = dict(a_kmajor=layout[0] == "n", b_kmajor=layout[1] == "t")
    outs = {}
    tol = 1e-5 if K <= 4096 else 3e-5        # fp32 accumulation over K products (unsplit kernel: 1.02e-5 at K = 8192)
    for split in (1, 0):
        lib().i2t_set_gemm_split_k(split)
        try:
            out = torch.full((M, N), 7.0, device=DEV)                      # stale contents must not leak into the result
            ops.gemm(A, Bm, bias=bias.to(DEV), out=out, **kw)
            assert relerr(out, ref + bias.double()) < tol, f"split={split}"
            outs[split] = out
            x = res.to(DEV).clone()
            ops.gemm(A, Bm, bias=bias.to(DEV), residual=x, out=x, **kw)    # x += a b^T + bias, in place (decode residual stream)
            assert relerr(x, ref + bias.double() + res.double()) < tol, f"split={split}"
            acc = res.to(DEV).clone()
            ops.gemm(A, Bm, out=acc, accumulate=True, **kw)                # dW += ...
            assert relerr(acc, ref + res.double()) < tol, f"split={split}"
            pitched = torch.zeros((M, N + 12), device=DEV)[:, :N]
            ops.gemm(A, Bm, out=pitched, **kw)
            assert relerr(pitched, ref) < tol and float(pitched.untyped_storage().nbytes()) > 0
        finally:
            lib().i2t_set_gemm_split_k(1)
    assert relerr(outs[1], outs[0].double()) < tol       # fp32 partial sums added in a different (and not fixed) order


def test_colsum():
    x = rnd(1000, 333, seed=13)
    out = torch.zeros(333, device=DEV)
    ops.colsum_(x.to(DEV), out)
    assert relerr(out, x.double().sum(0)) < 2e-6


# ------------------------------------------------------------------ attention ---------------------------------
def _ref_attention(qkv, B, T, H, mode, n_prompt):
    C = qkv.shape[1] // 3
    q, k, v = qkv.view(B, T, 3 * C).split(C, dim=2)
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    if mode == ops.MASK_NONE:
        vis = torch.ones(T, T, dtype=torch.bool)
    elif mode == ops.MASK_CAUSAL:
        vis = j <= i
    else:
        vis = (j <= i) & ((i < n_prompt) | (j >= n_prompt))
    mask = torch.zeros(T, T, dtype=qkv.dtype).masked_fill(~vis, -float("inf"))[None, None]
    y = O.sdpa(O.split_heads(q, H), O.split_heads(k, H), O.split_heads(v, H), mask)
    return O.merge_heads(y).reshape(B * T, C)


@pytest.mark.parametrize("B,T,H,hs,mode,n_prompt", [(2, 197, 12, 64, 0, 0), (2, 256, 12, 64, 2, 8), (3, 70, 4, 32, 1, 0),
                                                    (1, 64, 2, 64, 2, 4), (2, 48, 4, 32, 2, 4), (2, 272, 12, 64, 1, 0),
                                                    (1, 300, 4, 64, 0, 0), (3, 130, 3, 64, 2, 16),
                                                    # block-boundary cases of the tcgen05 kernels: one row, one row into the second /
                                                    # third 128-row block, one row short of a full block, the 384-key maximum of the forward
                                                    (1, 1, 1, 64, 1, 0), (2, 129, 2, 64, 1, 0), (1, 255, 1, 64, 2, 3),
                                                    (1, 257, 2, 64, 1, 0), (1, 384, 2, 64, 0, 0)])
def test_attention_fwd_bwd(B, T, H, hs, mode, n_prompt):
    C = H * hs
    qkv = rnd(B * T, 3 * C, seed=14)
    dout = rnd(B * T, C, seed=15)
    qr = qkv.double().requires_grad_(True)
    ref = _ref_attention(qr, B, T, H, mode, n_prompt)
    ref.backward(dout.double())
    out, lse = ops.attention_packed(qkv.to(DEV), B, T, H, mode, n_prompt, want_lse=True)
    assert relerr(out, ref) < 3e-6
    dqkv = ops.attention_packed_bwd(qkv.to(DEV), out, dout.to(DEV), lse, B, T, H, mode, n_prompt)
    assert relerr(dqkv, qr.grad) < 2e-5
    out16 = ops.attention_packed(qkv.to(DEV).to(torch.bfloat16), B, T, H, mode, n_prompt)
    assert relerr(out16.float(), ref) < 2e-2
    # bf16 backward on the tensor cores (attention_tc.cu) against the fp64 reference on the bf16-rounded inputs, and against
    # the fp32-math kernel on the same bf16 inputs (A/B switch)
    q16 = qkv.to(DEV).to(torch.bfloat16)
    d16 = dout.to(DEV).to(torch.bfloat16)
    qr16 = q16.double().cpu().requires_grad_(True)
    ref16 = _ref_attention(qr16, B, T, H, mode, n_prompt)
    ref16.backward(d16.double().cpu())
    o16, lse16 = ops.attention_packed(q16, B, T, H, mode, n_prompt, want_lse=True)
    g_tc = ops.attention_packed_bwd(q16, o16, d16, lse16, B, T, H, mode, n_prompt)
    assert relerr(g_tc.float(), qr16.grad) < 2e-2
    lib().i2t_set_tensor_core_attention(0)
    try:
        o_f, lse_f = ops.attention_packed(q16, B, T, H, mode, n_prompt, want_lse=True)
        g_f = ops.attention_packed_bwd(q16, o_f, d16, lse_f, B, T, H, mode, n_prompt)
    finally:
        lib().i2t_set_tensor_core_attention(1)
    assert relerr(g_tc.float(), g_f.float()) < 2e-2
    # forward A/B: tcgen05 kernel (default where eligible) vs the mma.sync kernel vs the fp32-math kernel, same bf16 inputs
    lib().i2t_set_tensor_core_attention(2)
    try:
        o_mma, lse_mma = ops.attention_packed(q16, B, T, H, mode, n_prompt, want_lse=True)
    finally:
        lib().i2t_set_tensor_core_attention(1)
    assert relerr(o16.float(), o_mma.float()) < 1e-2 and relerr(o16.float(), o_f.float()) < 1e-2
    assert float((lse16 - lse_mma).abs().max()) < 2e-2 and float((lse16 - lse_f).abs().max()) < 2e-2


@pytest.mark.parametrize("B,T,S,H,hs", [(2, 256, 8, 12, 64), (3, 20, 4, 4, 32), (1, 33, 16, 12, 64), (1, 7, 64, 2, 64)])
def test_xattn_fwd_bwd(B, T, S, H, hs):
    C = H * hs
    q, kv, dout = rnd(B * T, C, seed=16), rnd(B * S, 2 * C, seed=17), rnd(B * T, C, seed=18)
    qr, kvr = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    k, v = kvr.view(B, S, 2 * C).split(C, dim=2)
    ref = O.merge_heads(O.sdpa(O.split_heads(qr.view(B, T, C), H), O.split_heads(k, H), O.split_heads(v, H), None))
    ref = ref.reshape(B * T, C)
    ref.backward(dout.double())
    out = ops.xattn(q.to(DEV), kv.to(DEV), B, T, S, H)
    assert relerr(out, ref) < 3e-6
    dq, dkv = ops.xattn_bwd(q.to(DEV), kv.to(DEV), dout.to(DEV), B, T, S, H)
    assert relerr(dq, qr.grad) < 2e-5 and relerr(dkv, kvr.grad) < 2e-5
    # bf16: the tensor-core attention kernels with Tk = S (what XAttnFn runs), against fp64 on the bf16-rounded inputs
    q16, kv16, d16 = (t.to(DEV).to(torch.bfloat16) for t in (q, kv, dout))
    qr2, kvr2 = q16.double().cpu().requires_grad_(True), kv16.double().cpu().requires_grad_(True)
    k2, v2 = kvr2.view(B, S, 2 * C).split(C, dim=2)
    ref2 = O.merge_heads(O.sdpa(O.split_heads(qr2.view(B, T, C), H), O.split_heads(k2, H), O.split_heads(v2, H), None)).reshape(B * T, C)
    ref2.backward(d16.double().cpu())
    o16, lse16 = ops.xattn_tc(q16, kv16, B, T, S, H)
    assert relerr(o16.float(), ref2) < 1e-2
    dq16, dkv16 = ops.xattn_tc_bwd(q16, kv16, o16, d16, lse16, B, T, S, H)
    assert relerr(dq16.float(), qr2.grad) < 2e-2 and relerr(dkv16.float(), kvr2.grad) < 2e-2


# ------------------------------------------------------------------ encoder pieces ----------------------------
def test_patch_embed_matches_conv2d():
    B, img, p, d = 3, 64, 16, 96
    x = rnd(B, 3, img, img, seed=19)
    w, bias = rnd(d, 3, p, p, seed=20, scale=0.05), rnd(d, seed=21)
    cls, pos = rnd(1, 1, d, seed=22), rnd(1, (img // p) ** 2 + 1, d, seed=23)
    ref = F.conv2d(x.double(), w.double(), bias.double(), stride=p).reshape(B, d, -1).permute(0, 2, 1)
    ref = torch.cat([cls.double().expand(B, -1, -1), ref], 1) + pos.double()
    patches = ops.patch_im2col(x.to(DEV), p, torch.float32)
    po = ops.gemm(patches, w.view(d, -1).to(DEV), bias=bias.to(DEV))
    out = ops.vit_assemble(po, cls.to(DEV), pos.to(DEV), B)
    assert relerr(out, ref) < 3e-6


def test_lsh_tail_matches_oracle_indices_and_values():
    from image2text_b200.model_spec import synth_state_dict
    spec = O.default_spec(n_cls=3, lsh_num_bins=(4, 8, 20), lsh_num_proj=32, n_embd_out_vit=128, n_embd=128, vit_layers=0,
                          n_layer=0, vocab_size=8, block_size=8, tail="lsh", vit_image=32, gate_sizes=())
    sd = {k: v for k, v in synth_state_dict(spec).items() if ".lsh_emb." in k}
    feat = rnd(16, 768, seed=24)
    ref = O.lsh_tail(sd, spec, feat)
    d = {k: v.to(DEV) for k, v in sd.items()}
    keys = [(s, r) for s in range(3) for r in range(3)]
    tab = {n: torch.tensor([d[f"encoder.lsh_emb.{s}.emb.{r}.{leaf}"].data_ptr() for s, r in keys], dtype=torch.int64, device=DEV)
           for n, leaf in (("proj", "projection_mat"), ("grid", "grid"), ("emb", "emb.weight"))}
    nb = torch.tensor([4, 8, 20], dtype=torch.int32, device=DEV)
    out = torch.empty(16, 3, 128, device=DEV)
    idx = torch.empty(16, 3, 3, 32, dtype=torch.int32, device=DEV)
    featd = feat.to(DEV)
    call("i2t_lsh_tail", ptr(featd), ptr(tab["proj"]), ptr(tab["grid"]), ptr(tab["emb"]), ptr(nb), ptr(out), ptr(idx),
         16, 768, 3, 3, 32, 128, stream())
    for s in range(3):
        for r, nbins in enumerate((4, 8, 20)):
            kp = f"encoder.lsh_emb.{s}.emb.{r}."
            want = O.lsh_bucket_indices(feat, sd[kp + "projection_mat"], sd[kp + "grid"], nbins)
            assert torch.equal(idx[:, s, r].cpu().long(), want), (s, r)     # integer hashing: bit-exact
    assert relerr(out, ref) < 2e-6


def test_lsh_tail_backward_is_the_embedding_bag_mean_scatter():
    """d emb[idx[b,s,r,p]] += d out[b,s] / n_proj (F.embedding_bag(mode='mean') backward, models/layers.py:139-144), against the
    CPU autograd of the same lookups; second call: the kernel ADDS to what the buffers hold (gradient accumulation)."""
    import ctypes
    B, n_cls, n_res, n_proj, E, bins = 5, 3, 2, 8, 96, (4, 9)
    g = torch.Generator().manual_seed(31)
    dout = rnd(B, n_cls, E, seed=32)
    rows = [(bins[r] + 1) * n_proj for r in range(n_res)]
    idx = torch.stack([torch.stack([torch.stack([torch.randint(0, bins[r] + 1, (n_proj,), generator=g) +
                                                 (bins[r] + 1) * torch.arange(n_proj) for r in range(n_res)])
                                    for _ in range(n_cls)]) for _ in range(B)])                     # (B, n_cls, n_res, n_proj)
    tabs = [torch.zeros(rows[r], E, requires_grad=True) for _ in range(n_cls) for r in range(n_res)]
    out = torch.stack([sum(torch.nn.functional.embedding_bag(idx[:, s, r], tabs[s * n_res + r], mode="mean")
                           for r in range(n_res)) for s in range(n_cls)], dim=1)
    (out * dout).sum().backward()
    dev = [torch.zeros(rows[r], E, device=DEV) for _ in range(n_cls) for r in range(n_res)]
    host = (ctypes.c_void_p * len(dev))(*[t.data_ptr() for t in dev])
    doutd, idxd = dout.to(DEV), idx.to(DEV).int().contiguous()
    for rep in (1, 2):
        call("i2t_lsh_tail_bwd", ptr(doutd), ptr(idxd), ctypes.addressof(host), B, n_cls, n_res, n_proj, E, stream())
        for t, ref in zip(dev, tabs):
            assert relerr(t, rep * ref.grad) < 2e-6


def test_embed_prompt_concat():
    B, S, n, C, V, blk = 3, 20, 4, 128, 97, 22
    ids = torch.randint(0, V, (B, S), generator=torch.Generator().manual_seed(25))
    prompt, wte, wpe = rnd(B, n, C, seed=26), rnd(V, C, seed=27), rnd(blk, C, seed=28)
    T = min(n + S, blk)
    ref = torch.cat([prompt, wte[ids]], 1)[:, :T] + wpe[:T]
    out = ops.embed(ids.to(DEV), prompt.to(DEV), wte.to(DEV), wpe.to(DEV), B, T, n, S)
    assert torch.equal(out.cpu(), ref)


# ------------------------------------------------------------------ decode kernels ----------------------------
@pytest.mark.parametrize("B,N,K,wdt", [(8, 2304, 768, torch.float32), (8, 768, 3072, torch.float32), (3, 50257, 768, torch.float32),
                                       (16, 384, 128, torch.float32), (8, 2304, 768, torch.bfloat16), (5, 1000, 3072, torch.bfloat16)])
def test_dec_linear(B, N, K, wdt):
    x = rnd(B, K, seed=29) + 0.2
    g, be = 1 + 0.1 * rnd(K, seed=30), 0.1 * rnd(K, seed=31)
    w = rnd(N, K, seed=32, scale=0.05).to(wdt)
    bias, res = rnd(N, seed=33), rnd(B, N, seed=34)
    wcode = ops.F32 if wdt == torch.float32 else ops.BF16
    xn = F.layer_norm(x.double(), (K,), g.double(), be.double(), 1e-5)
    if wdt == torch.bfloat16:
        xn = xn.float().to(torch.bfloat16).double()
    ref = O.gelu_tanh(xn @ w.double().t() + bias.double()) + res.double()
    out = torch.empty(B, N, device=DEV)
    # device copies are bound to names: a temporary would be recycled by the caching allocator before the launch
    xd, gd, bed, wdev, biasd, resd = (t.to(DEV) for t in (x, g, be, w, bias, res))
    call("i2t_dec_linear", ptr(xd), ptr(gd), ptr(bed), 1e-5, ptr(wdev), ptr(biasd), ptr(resd), ptr(out), N, B, N, K,
         ops.ACT_GELU_TANH, wcode, 0, None, None, 0, 0, 0, None, stream())
    assert relerr(out, ref) < (5e-6 if wdt == torch.float32 else 3e-3)
    # no LayerNorm, no bias, in-place residual
    xin = x.double() if wdt == torch.float32 else x.to(torch.bfloat16).double()
    ref2 = xin @ w.double().t() + res.double()
    acc = resd.clone()
    call("i2t_dec_linear", ptr(xd), None, None, 1e-5, ptr(wdev), None, ptr(acc), ptr(acc), N, B, N, K, 0, wcode, 0,
         None, None, 0, 0, 0, None, stream())
    assert relerr(acc, ref2) < (5e-6 if wdt == torch.float32 else 3e-3)


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_dec_qkv_append_and_attention(cdt):
    B, H, hs, Tmax, pos = 4, 12, 64, 40, 17
    C = H * hs
    code = ops.F32 if cdt == torch.float32 else ops.BF16
    x = rnd(B, C, seed=35)
    w, bias = rnd(3 * C, C, seed=36, scale=0.05).to(cdt), rnd(3 * C, seed=37)
    kc = rnd(B, Tmax, C, seed=38).to(cdt).to(DEV)
    vc = rnd(B, Tmax, C, seed=39).to(cdt).to(DEV)
    kc0, vc0 = kc.clone(), vc.clone()
    q = torch.empty(B, C, device=DEV)
    posd = torch.tensor([pos], dtype=torch.int32, device=DEV)
    xd, wdev, biasd = x.to(DEV), w.to(DEV), bias.to(DEV)
    call("i2t_dec_linear", ptr(xd), None, None, 1e-5, ptr(wdev), ptr(biasd), None, ptr(q), C, B, 3 * C, C, 0,
         code, 1, ptr(kc), ptr(vc), Tmax * C, C, code, ptr(posd), stream())
    xin = x.double() if cdt == torch.float32 else x.to(torch.bfloat16).double()
    qkv = xin @ w.double().t() + bias.double()
    tol = 3e-6 if cdt == torch.float32 else 1e-2
    assert relerr(q, qkv[:, :C]) < tol
    assert relerr(kc[:, pos].float(), qkv[:, C:2 * C]) < tol and relerr(vc[:, pos].float(), qkv[:, 2 * C:]) < tol
    keep = torch.ones(Tmax, dtype=torch.bool)
    keep[pos] = False
    assert torch.equal(kc[:, keep], kc0[:, keep]) and torch.equal(vc[:, keep], vc0[:, keep])   # only slot `pos` written
    y = torch.empty(B, C, device=DEV)
    call("i2t_dec_attn", ptr(q), C, ptr(kc), ptr(vc), Tmax * C, C, ptr(y), C, ptr(posd), 1, B, H, hs, code, stream())
    qq = q.cpu().double() if cdt == torch.float32 else q.cpu().to(torch.bfloat16).double()
    kk, vv = kc[:, :pos + 1].cpu().double(), vc[:, :pos + 1].cpu().double()
    ref = O.merge_heads(O.sdpa(O.split_heads(qq[:, None], H), O.split_heads(kk, H), O.split_heads(vv, H), None))[:, 0]
    assert relerr(y, ref) < (3e-6 if cdt == torch.float32 else 2e-3)


# ------------------------------------------------------------------ sampler -----------------------------------
def _run_sampler(logits, ids, cur_len, temperature, top_k, ngrams, seed=1, want_probs=True, nucleus_p=0.0):
    B, V = logits.shape
    lg = logits.to(DEV).clone()
    idd = ids.to(DEV).clone()
    ng = torch.tensor(list(ngrams) or [0], dtype=torch.int32, device=DEV)
    probs = torch.empty(B, V, device=DEV) if want_probs else None
    call("i2t_sample", ptr(lg), V, B, V, ptr(idd), idd.shape[1], None, 0, cur_len, temperature, top_k or 0, float(nucleus_p),
         ptr(ng), len(ngrams), seed, None, ptr(probs), None, 1, stream())
    return idd[:, cur_len].cpu(), (probs.cpu() if want_probs else None)


@pytest.mark.parametrize("V,top_k,p", [(613, None, 0.9), (613, 40, 0.5), (50257, None, 0.95), (50257, 64, 0.3), (50257, None, 0.01)])
def test_sampler_nucleus_matches_oracle(V, top_k, p):
    """top-p filter (reference models/vision_encoder_decoder.py:160-172, restated in oracle.nucleus_filter): the device
    sampler's distribution equals the reference's sorted / cumsum / renormalised one scattered back to token order."""
    B = 4
    logits = rnd(B, V, seed=70) * 3.0
    ids = torch.zeros(B, 9, dtype=torch.int64)
    ids[:, :8] = torch.randint(0, V, (B, 8), generator=torch.Generator().manual_seed(71))
    spec = dict(no_repeat_n_grams=(2, 3, 4, 5))
    base = O.next_token_probs(logits, ids[:, :8], spec, 0.8, top_k)
    sp, si = O.nucleus_filter(base, p)
    want = torch.zeros_like(base).scatter_(1, si, sp)
    tok, probs = _run_sampler(logits, ids, 8, 0.8, top_k, (2, 3, 4, 5), nucleus_p=p)
    assert float((probs - want).abs().max()) < 2e-6, float((probs - want).abs().max())
    assert int((probs > 0).sum()) == int((want > 0).sum())
    assert bool((want.gather(1, tok[:, None]) > 0).all())
    if p == 0.01:                                       # p below the top probability: exactly the arg-max survives
        assert torch.equal(tok, base.argmax(-1))


def test_sampler_ngram_ban_topk_softmax_match_oracle(golden):
    g = golden("ngram")
    ids = torch.from_numpy(g["ids"])
    B, L = ids.shape
    V = 16
    logits = torch.from_numpy(g["scores"])
    spec = dict(no_repeat_n_grams=(2, 3, 4, 5))
    buf = torch.zeros(B, L + 1, dtype=torch.int64)
    buf[:, :L] = ids
    for cur in (1, 2, 3, 4, 5, 9, L):
        for top_k, temp in ((None, 1.0), (3, 0.7), (1, 1.0)):
            want = O.next_token_probs(logits, ids[:, :cur], spec, temp, top_k)
            tok, probs = _run_sampler(logits, buf, cur, temp, top_k, (2, 3, 4, 5))
            assert float((probs - want).abs().max()) < 1e-6, (cur, top_k)
            banned = torch.from_numpy(g["banned_full"] if cur == L else g[f"banned_len{cur}"])
            assert bool((probs[banned] == 0).all())
            assert bool((want.gather(1, tok[:, None]) > 0).all())          # never samples a dropped token
            if top_k == 1:
                assert torch.equal(tok, want.argmax(-1))


def test_sampler_greedy_is_argmax_full_vocab():
    B, V = 8, 50257
    logits = rnd(B, V, seed=40)
    ids = torch.zeros(B, 4, dtype=torch.int64)
    tok, _ = _run_sampler(logits, ids, 1, 1.0, 1, (), want_probs=False)
    assert torch.equal(tok, logits.argmax(-1))
    tok, probs = _run_sampler(logits, ids, 1, 0.8, 16, ())
    want = O.next_token_probs(logits, ids[:, :1], dict(no_repeat_n_grams=()), 0.8, 16)
    assert float((probs - want).abs().max()) < 1e-6
    assert bool(((probs > 0).sum(-1) == 16).all())


def test_sampler_greedy_fast_path_equals_general_sampler():
    """top_k = 1 takes the one-pass arg-max kernel: same picks as the general sampler (A/B switch) and as the oracle's
    ban + arg-max, with n-gram bans active, an unaligned row pitch, and the lowest id winning an exact tie."""
    B, V = 37, 50257
    logits = rnd(B, V, seed=44)
    g = torch.Generator().manual_seed(45)
    ids = torch.randint(0, 6, (B, 24), generator=g)              # tiny alphabet: every n-gram repeats -> many bans
    logits[:, :6] += 6.0                                          # ... and the banned tokens would otherwise win
    logits[3, 100] = logits[3, 20000] = 50.0                      # exact tie at the top
    cur = 20
    picks = {}
    for fast in (1, 0):
        lib().i2t_set_sampler_greedy_fast_path(fast)
        try:
            picks[fast], _ = _run_sampler(logits, ids, cur, 0.7, 1, (2, 3, 4, 5), want_probs=False)
        finally:
            lib().i2t_set_sampler_greedy_fast_path(1)
    banned = O.apply_ngram_ban(ids[:, :cur], logits.clone(), (2, 3, 4, 5))
    want = banned.argmax(-1)
    assert torch.equal(picks[1], want) and picks[1][3] == 100
    rows = [i for i in range(B) if i != 3]                         # the general sampler draws among exact ties
    assert torch.equal(picks[0][rows], want[rows])
    # pitched rows (GEMM-mode decode: 16-byte pitch) and a device-side position / ticket
    pitched = torch.zeros(B, V + 3, device=DEV)
    pitched[:, :V] = logits.to(DEV)
    idd = ids.to(DEV).clone()
    pos = torch.tensor([cur - 1], dtype=torch.int32, device=DEV)
    ticket = torch.zeros(1, dtype=torch.int32, device=DEV)
    ng = torch.tensor([2, 3, 4, 5], dtype=torch.int32, device=DEV)
    call("i2t_sample", ptr(pitched), V + 3, B, V, ptr(idd), idd.shape[1], ptr(pos), 1, 0, 1.0, 1, 0.0, ptr(ng), 4, 1, None, None,
         ptr(ticket), 1, stream())
    assert torch.equal(idd[:, cur].cpu(), want) and int(pos.item()) == cur and int(ticket.item()) == 0


def test_sampler_distribution_matches_topk_softmax():
    """top-k sampling must follow the reference's token distribution: chi-square over 20000 draws."""
    V, k, n = 1000, 8, 20000
    logits = rnd(1, V, seed=41).repeat(n, 1)
    ids = torch.zeros(n, 2, dtype=torch.int64)
    tok, _ = _run_sampler(logits, ids, 1, 1.0, k, (), seed=1234, want_probs=False)
    want = O.next_token_probs(logits[:1], ids[:1, :1], dict(no_repeat_n_grams=()), 1.0, k)[0]
    support = want.nonzero().flatten()
    counts = torch.bincount(tok, minlength=V).double()
    assert counts[want == 0].sum() == 0
    exp = want[support].double() * n
    chi2 = float(((counts[support] - exp) ** 2 / exp).sum())
    assert chi2 < 30.0, chi2          # 7 dof: P(chi2 > 30) ~ 1e-4
    tok2, _ = _run_sampler(logits, ids, 1, 1.0, k, (), seed=99, want_probs=False)
    assert not torch.equal(tok, tok2)


# ------------------------------------------------------------------ loss / optimisers -------------------------
@pytest.mark.parametrize("distill", [False, True])
def test_lm_loss_and_gradient(distill):
    from image2text_b200.autograd_ops import LmLossFn
    from image2text_b200.synthetic import synth_labels
    B, T, V = 3, 24, 1031
    logits = rnd(B, T, V, seed=42, scale=2.0)
    teacher = rnd(B, T, V, seed=43, scale=2.0) if distill else None
    labels = synth_labels(B, 32, V, seed=44, min_len=3, max_len=20, eos=V - 1)
    kw = dict(temperature=1.3, alpha=0.4, weight_fn="inverse_sqrt_position", eos_token_weight=2.0, eos_token_id=V - 1) \
        if distill else {}
    lr = logits.double().requires_grad_(True)
    ref = O.lm_loss(lr, labels, teacher.double() if distill else None, **kw)
    ref.backward()
    lg = logits.to(DEV).requires_grad_(True)
    loss = LmLossFn.apply(lg, teacher.to(DEV) if distill else None, labels.to(DEV), kw.get("temperature", 1.0),
                          kw.get("alpha"), kw.get("weight_fn", "constant"), kw.get("eos_token_weight"), V - 1, -100)
    (loss * 0.5).backward()
    assert abs(float(loss) - float(ref)) < 2e-6 * abs(float(ref))
    assert relerr(lg.grad, 0.5 * lr.grad) < 1e-5


def _tables(tensors, chunk=4096):
    n = len(tensors[0])
    table = torch.tensor([[t[i].data_ptr() if t[i] is not None else 0 for t in tensors] for i in range(n)], dtype=torch.int64)
    ct, co, cl = [], [], []
    for i in range(n):
        numel = tensors[0][i].numel()
        for off in range(0, numel, chunk):
            ct.append(i); co.append(off); cl.append(min(chunk, numel - off))
    dev = lambda x, dt: torch.tensor(x, dtype=dt, device=DEV)
    return table.to(DEV), dev(ct, torch.int32), dev(co, torch.int64), dev(cl, torch.int32), len(ct)


@pytest.mark.parametrize("name,fn", [("adamw", "i2t_adamw_multi"), ("adamw_nowd", "i2t_adamw_multi"), ("snradam", "i2t_snradam_multi")])
def test_fused_optimizers_match_reference_steps(golden, name, fn):
    g = golden("optim")
    hp = dict(adamw=(3e-3, 0.9, 0.95, 0.1), adamw_nowd=(1e-3, 0.9, 0.999, 0.0), snradam=(3e-3, 0.9, 0.95, 0.1))[name]
    p = torch.from_numpy(g[f"{name}_p0"]).to(DEV).clone()
    extra = rnd(7, 5, seed=45).to(DEV)          # a second, odd-sized, unaligned-tail tensor in the same launch
    extra0 = extra.clone()
    ms = [torch.zeros_like(p), torch.zeros_like(extra)]
    vs = [torch.zeros_like(p), torch.zeros_like(extra)]
    grads = [torch.empty_like(p), torch.zeros_like(extra)]
    table, ct, co, cl, nch = _tables(([p, extra], grads, ms, vs), chunk=512)
    for step in range(4):
        grads[0].copy_(torch.from_numpy(g[f"{name}_g{step}"]))
        call(fn, ptr(table), ptr(ct), ptr(co), ptr(cl), nch, hp[0], hp[1], hp[2], 1e-8, hp[3], step + 1, 1.0, stream())
        assert relerr(p, torch.from_numpy(g[f"{name}_p{step + 1}"])) < 2e-6, step
    decay = (1 - hp[0] * hp[3]) ** 4
    assert relerr(extra, extra0 * decay) < 1e-6          # zero gradient: only weight decay acts


def test_ema_multi_and_actfn_and_gradnorm():
    pm, p = rnd(1000, 37, seed=46).to(DEV), rnd(1000, 37, seed=47).to(DEV)
    want = pm * 0.995 + p * (1 - 0.995)
    table, ct, co, cl, nch = _tables(([pm], [p], [None], [None]), chunk=4096)
    call("i2t_ema_multi", ptr(table), ptr(ct), ptr(co), ptr(cl), nch, 0.995, stream())
    assert relerr(pm, want) < 1e-6
    g = rnd(8, 256, 768, seed=48).to(DEV)
    out = torch.empty_like(g)
    acc = torch.empty(1, dtype=torch.float64, device=DEV)
    call("i2t_gradnorm_scale", ptr(g), ptr(out), ptr(acc), g.numel(), ops.F32, stream())
    assert relerr(out, g.double() / (g.double().norm() + 1e-6)) < 2e-6
    for act, f in ((ops.ACT_GELU_TANH, O.gelu_tanh), (ops.ACT_GELU_ERF, O.gelu_erf)):
        z = rnd(64, 512, seed=49, scale=2.0)
        zr = z.double().requires_grad_(True)
        f(zr).backward(torch.ones_like(zr) * 0.7)
        h = torch.empty(64, 512, device=DEV)
        dz = torch.empty(64, 512, device=DEV)
        dh = torch.full((64, 512), 0.7, device=DEV)
        zd = z.to(DEV)
        call("i2t_act_fwd", ptr(zd), ptr(h), z.numel(), act, ops.F32, ops.F32, stream())
        call("i2t_act_bwd", ptr(zd), ptr(dh), ptr(dz), z.numel(), act, ops.F32, ops.F32, stream())
        assert relerr(h, f(z.double())) < 2e-6 and relerr(dz, zr.grad) < 5e-6
