"""ctypes binding of include/b200_decoder.h.  There is deliberately no fallback: if the shared
library is missing or the device is not an sm_100 part, importing callers get a RuntimeError."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200decoder.so")

_lib = None


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("D", C.c_void_p), ("ldd", C.c_int64), ("d_fp32", C.c_int32), ("accumulate", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("relu_mask", C.c_void_p), ("ldm", C.c_int64),
        ("act", C.c_int32), ("split_k", C.c_int32), ("block_n", C.c_int32),
    ]


class AttnFwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_bs", C.c_int64), ("q_ts", C.c_int64),
        ("k", C.c_void_p), ("k_bs", C.c_int64), ("k_ts", C.c_int64),
        ("v", C.c_void_p), ("v_bs", C.c_int64), ("v_ts", C.c_int64),
        ("o", C.c_void_p), ("o_bs", C.c_int64), ("o_ts", C.c_int64),
        ("lse", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32), ("hd", C.c_int32),
        ("causal", C.c_int32),
        ("key_tokens", C.c_void_p), ("pad_idx", C.c_int64),
        ("key_pad_mask", C.c_void_p),
        ("scale", C.c_float),
        ("cu_q", C.c_void_p), ("cu_k", C.c_void_p),
        ("total_q", C.c_int32), ("total_k", C.c_int32),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("f", AttnFwdArgs),
        ("d_o", C.c_void_p), ("do_bs", C.c_int64), ("do_ts", C.c_int64),
        ("dq", C.c_void_p), ("dq_bs", C.c_int64), ("dq_ts", C.c_int64),
        ("dk", C.c_void_p), ("dk_bs", C.c_int64), ("dk_ts", C.c_int64),
        ("dv", C.c_void_p), ("dv_bs", C.c_int64), ("dv_ts", C.c_int64),
    ]


class EngineConfig(C.Structure):
    _fields_ = [
        ("vocab_size", C.c_int32), ("embed_dim", C.c_int32), ("num_heads", C.c_int32),
        ("num_layers", C.c_int32), ("ff_dim", C.c_int32), ("max_seq_len", C.c_int32),
        ("enc_dim", C.c_int32),
        ("pad_idx", C.c_int64),
        ("ln_eps", C.c_float),
        ("act", C.c_int32),
    ]


_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/b200_decoder.h one to one (checked by
# tests/test_abi.py against the header text and the built library's export table).
SIGNATURES = {
    "b200_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_check_device": (C.c_int, [C.c_int]),
    "b200_launch_count": (C.c_longlong, []),
    "b200_gemm_profile_begin": (C.c_int, [_I32]),
    "b200_gemm_profile_end": (C.c_int, [_P, _P, _P, _P, _P, _I32]),
    "b200_gemm": (C.c_int, [C.POINTER(GemmArgs), _P]),
    "b200_gemm_check": (C.c_int, [C.POINTER(GemmArgs), _P]),
    "b200_lmhead_ce_fwd": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _I32, _I32, _I32, _I64, _P, _P, _P, _P, _P, _P]),
    "b200_lmhead_ce_bwd": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _I32, _I32, _I32, _I64, _P, _P, _P, _I64, _P]),
    "b200_lmhead_argmax": (C.c_int, [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "b200_embed_pe_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _P]),
    "b200_embed_bwd": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I64, _F, _P]),
    "b200_layernorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _F, _P]),
    "b200_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _P]),
    "b200_colsum": (C.c_int, [_P, _I64, _P, _I32, _I32, _P]),
    "b200_cast_f32_to_bf16": (C.c_int, [_P, _P, _I64, _P]),
    "b200_cast_bf16_to_f32": (C.c_int, [_P, _P, _I64, _P]),
    "b200_attn_fwd": (C.c_int, [C.POINTER(AttnFwdArgs), _P]),
    "b200_attn_bwd": (C.c_int, [C.POINTER(AttnBwdArgs), _P]),
    "b200_attn_tc_trace": (C.c_int, [_P]),
    "b200_grad_sumsq": (C.c_int, [_P, _I64, _P, _P]),
    "b200_adamw_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _F, _F, _I32, _P]),
    "b200_adamw_step_dev": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _F, _P, _F, _F, _F, _F, _P, _P]),
    "b200_engine_create": (C.c_int, [C.POINTER(EngineConfig), C.POINTER(C.c_void_p)]),
    "b200_engine_destroy": (None, [_P]),
    "b200_engine_param_count": (_I64, [_P]),
    "b200_engine_param_offset": (_I64, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "b200_engine_param_name": (C.c_int, [_P, _I32, C.c_char_p, _I32]),
    "b200_engine_num_params": (_I32, [_P]),
    "b200_engine_bind": (C.c_int, [_P, _P, _P, _P, _P]),
    "b200_engine_set_memory_dtype": (C.c_int, [_P, _I32]),
    "b200_engine_set_dropout": (C.c_int, [_P, _F, _P]),
    "b200_engine_workspace_bytes": (_I64, [_P, _I32, _I32, _I32, _I32, _I32]),
    "b200_engine_set_workspace": (C.c_int, [_P, _P, _I64]),
    "b200_engine_forward_logits": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "b200_engine_forward_loss": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I64, _I32, _P, _P]),
    "b200_engine_forward_loss_packed": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I64, _I32, _P, _I32, _P, _P]),
    "b200_engine_backward": (C.c_int, [_P, _P, _P, _P, _I32, _P]),
    "b200_engine_backward_parts": (C.c_int, [_P, _P, _P, _I32, _I32, _P]),
    "b200_engine_backward_from_dlogits": (C.c_int, [_P, _P, _P, _P]),
    "b200_engine_grad_buckets": (_I32, [_P, _P, _P, _I32]),
    "b200_engine_decode_workspace_bytes": (_I64, [_P, _I32, _I32, _I32, _I32, _I32]),
    "b200_engine_decode_plan_info": (C.c_int, [_P, _P, _I32]),
    "b200_engine_decode_begin": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _I64, _P]),
    "b200_engine_decode_step": (C.c_int, [_P, _P, _I32, _P, _P]),
    "b200_engine_generate_greedy": (C.c_int, [_P, _I64, _I64, _I32, _I32, _P, _P, _P]),
    "b200_engine_generate_beam": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _P, _P]),
}


def lib():
    """Load (once) and return the CDLL; raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "b200 decoder library not built: %s is missing (run `python __graft_entry__.py`); "
                "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)   # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc, what="b200 call"):
    if rc != 0:
        msg = lib().b200_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
