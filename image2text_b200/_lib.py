"""ctypes binding of libi2t.so (include/i2t.h).  There is no fallback: if the library is missing or a call
fails, the caller gets an exception -- the CUDA path is the only path."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libi2t.so")

P, I, L, F, D, U64 = c_void_p, c_int, c_int64, c_float, c_double, c_uint64

# name -> (restype, argtypes); every symbol include/i2t.h declares
SIGNATURES = {
    "i2t_version": (c_int, []),
    "i2t_last_error": (c_char_p, []),
    "i2t_launch_count": (c_int64, []),
    "i2t_set_tensor_core_gemm": (None, [I]),
    "i2t_set_pdl": (None, [I]),
    "i2t_set_tensor_core_attention": (None, [I]),
    "i2t_set_gemm_cta_pair": (None, [I]),
    "i2t_set_gemm_tma_store": (None, [I]),
    "i2t_set_gemm_split_k": (None, [I]),
    "i2t_set_sampler_greedy_fast_path": (None, [I]),
    "i2t_layernorm_fwd": (c_int, [P, P, P, P, P, P, L, L, L, F, I, I, P]),
    "i2t_layernorm_bwd": (c_int, [P, P, P, P, P, P, P, P, L, L, I, I, I, P]),
    "i2t_layernorm_bwd_add": (c_int, [P, P, P, P, P, P, P, P, P, L, L, I, I, I, P]),
    "i2t_gemm": (c_int, [P, P, P, P, P, L, L, L, L, L, L, I, I, I, I, I, I, I, P]),
    "i2t_gemm_ex": (c_int, [P, P, P, P, P, L, L, L, L, L, L, I, I, I, I, I, I, I, I, P]),
    "i2t_colsum": (c_int, [P, P, L, L, L, I, P]),
    "i2t_attn_fwd": (c_int, [P, P, P, P, P, L, L, L, L, L, L, L, L, L, I, L, I, I, P]),
    "i2t_attn_bwd_workspace_bytes": (c_int64, [L, L, L, L]),
    "i2t_attn_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, L, L, L, L, L, L, L, L, L, I, L, I, P]),
    "i2t_patch_im2col": (c_int, [P, P, L, L, L, L, I, P]),
    "i2t_vit_assemble": (c_int, [P, P, P, P, L, L, L, I, P]),
    "i2t_lsh_tail": (c_int, [P, P, P, P, P, P, P, L, L, L, L, L, L, P]),
    "i2t_attn_bwd_trace": (c_int, [P]),
    "i2t_lsh_tail_bwd": (c_int, [P, P, P, L, L, L, L, L, P]),
    "i2t_peer_lookup_fwd": (c_int, [P, P, P, P, P, P, P, P, P, P, P, L, L, L, L, L, L, P]),
    "i2t_peer_lookup_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, P, P, L, L, L, L, L, L, P]),
    "i2t_embed_fwd": (c_int, [P, P, P, P, P, L, L, L, L, L, P]),
    "i2t_xattn_fwd": (c_int, [P, P, P, P, L, L, L, L, L, L, L, L, I, I, P]),
    "i2t_xattn_bwd": (c_int, [P, P, P, P, P, P, P, L, L, L, L, L, L, L, L, L, L, I, P]),
    "i2t_dec_embed": (c_int, [P, P, P, P, P, L, L, L, L, P]),
    "i2t_dec_advance": (c_int, [P, P]),
    "i2t_dec_embed_rows": (c_int, [P, L, P, P, P, L, L, P]),
    "i2t_dec_kv_append": (c_int, [P, L, P, P, L, L, L, I, P, P]),
    "i2t_dec_linear": (c_int, [P, P, P, F, P, P, P, P, L, L, L, L, I, I, I, P, P, L, L, I, P, P]),
    "i2t_dec_attn": (c_int, [P, L, P, P, L, L, P, L, P, L, L, L, L, I, P]),
    "i2t_dec_attn_append": (c_int, [P, L, P, P, L, L, P, L, P, L, L, L, L, I, P, P, L, I, P]),
    "i2t_dec_layernorm": (c_int, [P, P, P, P, L, L, F, I, P, L, P]),
    "i2t_dec_act": (c_int, [P, P, L, I, I, P]),
    "i2t_sample": (c_int, [P, L, L, L, P, L, P, I, L, F, L, F, P, L, U64, P, P, P, I, P]),
    "i2t_decode_mega": (c_int, [P, P, P, L, L, L, L, L, L, L, I, P, L, P, P, P, P, P, P, F, L, P, L, P, P, L, P, P]),
    "i2t_decode_mega2_max_keys": (c_int, []),
    "i2t_decode_mega2": (c_int, [P, P, P, L, P, L, L, L, L, L, L, L, L, L, L, P, L, P, P, P, P, P, P, P, F, L, P, L, P, L, L, P, P]),
    "i2t_decode_mega3_max_keys": (c_int, []),
    "i2t_set_decode_poll_sleep": (None, [I]),
    "i2t_decode_mega3_grid": (c_int, []),
    "i2t_decode_mega3_tile_bytes": (c_int64, [L]),
    "i2t_decode_mega3_pack": (c_int, [P, L, L, L, P, P, L, P]),
    "i2t_decode_mega3_prepare": (c_int, [P, L, P, P, L, L, L, L, L, P, L, L, L, L, P, L, P, P]),
    "i2t_decode_mega3": (c_int, [P, P, P, P, L, L, L, L, L, L, L, L, L, L, L, P, L, P, P, L, P, P, P, P, P, L, F, L, P, L, P, L,
                                 L, P, L, L, P]),
    "i2t_act_fwd": (c_int, [P, P, L, I, I, I, P]),
    "i2t_act_bwd": (c_int, [P, P, P, L, I, I, I, P]),
    "i2t_embed_bwd": (c_int, [P, P, P, L, L, L, L, L, P]),
    "i2t_gradnorm_scale": (c_int, [P, P, P, L, I, P]),
    "i2t_lm_loss": (c_int, [P, P, P, P, P, P, P, L, L, L, L, L, F, F, I, I, F, L, L, L, L, F, I, P]),
    "i2t_contrastive_loss": (c_int, [P, P, P, P, P, P, L, L, L, F, I, I, F, L, L, F, P]),
    "i2t_scale_inplace": (c_int, [P, P, L, I, P]),
    "i2t_l2norm_fwd": (c_int, [P, P, L, L, F, P]),
    "i2t_l2norm_bwd": (c_int, [P, P, P, L, L, F, P]),
    "i2t_attn_fwd_dropout": (c_int, [P, P, P, P, P, L, L, L, L, L, L, L, L, L, I, L, I, I, F, P, L, P]),
    "i2t_attn_bwd_dropout": (c_int, [P, P, P, P, P, P, P, P, P, P, L, L, L, L, L, L, L, L, L, I, L, I, F, P, L, P]),
    "i2t_attn_bwd_dropout_tok": (c_int, [P, P, P, P, P, P, P, P, P, P, L, L, L, L, L, L, I, L, I, F, P, L, F, L, P]),
    "i2t_dropout_add_fwd": (c_int, [P, P, P, L, F, P, L, I, P]),
    "i2t_dropout_bwd": (c_int, [P, P, L, F, P, L, I, I, P]),
    "i2t_token_dropout": (c_int, [P, L, L, L, L, F, P, L, I, P]),
    "i2t_rng_advance": (c_int, [P, P]),
    "i2t_adamw_multi": (c_int, [P, P, P, P, L, D, D, D, D, D, L, D, P]),
    "i2t_snradam_multi": (c_int, [P, P, P, P, L, D, D, D, D, D, L, D, P]),
    "i2t_ema_multi": (c_int, [P, P, P, P, L, D, P]),
    "i2t_cast_bf16_multi": (c_int, [P, P, P, P, L, P]),
}

_lib = None


class I2TError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise I2TError(f"{LIB_PATH} is missing: run `python -m image2text_b200.build` (there is no CPU / eager "
                           f"fallback for the hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if os.environ.get("I2T_PDL") is not None:          # A/B switch for measurements
            handle.i2t_set_pdl(int(os.environ["I2T_PDL"]))
        if os.environ.get("I2T_TC_GEMM") is not None:
            handle.i2t_set_tensor_core_gemm(int(os.environ["I2T_TC_GEMM"]))
        if os.environ.get("I2T_GEMM_TMA_STORE") is not None:
            handle.i2t_set_gemm_tma_store(int(os.environ["I2T_GEMM_TMA_STORE"]))
        if os.environ.get("I2T_GEMM_SPLITK") is not None:
            handle.i2t_set_gemm_split_k(int(os.environ["I2T_GEMM_SPLITK"]))
        if os.environ.get("I2T_GEMM_PAIR") is not None:
            handle.i2t_set_gemm_cta_pair(int(os.environ["I2T_GEMM_PAIR"]))
        if os.environ.get("I2T_POLL_SLEEP") is not None:
            handle.i2t_set_decode_poll_sleep(int(os.environ["I2T_POLL_SLEEP"]))
        if os.environ.get("I2T_TC_ATTN") is not None:
            handle.i2t_set_tensor_core_attention(int(os.environ["I2T_TC_ATTN"]))
        _lib = handle
    return _lib


def call(name: str, *args):
    """Invoke a status-returning entry point; raise with the library's message on failure."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        msg = lib().i2t_last_error()
        raise I2TError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().i2t_launch_count())
