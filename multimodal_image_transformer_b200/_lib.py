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


def lib():
    """Load (once) and return the CDLL; raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "b200 decoder library not built: %s is missing (run `python __graft_entry__.py`); "
                "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.b200_last_error.restype = C.c_char_p
        for name in ("b200_engine_param_count", "b200_engine_param_offset",
                     "b200_engine_workspace_bytes", "b200_engine_decode_workspace_bytes"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_int64
    return _lib


def check(rc, what="b200 call"):
    if rc != 0:
        msg = lib().b200_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
