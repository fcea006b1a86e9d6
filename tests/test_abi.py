"""The C ABI: every symbol include/b200_decoder.h declares is exported by the built library and
bound (with a signature) by the ctypes loader — no compute calls, runs without a GPU."""
import os
import re
import subprocess

from multimodal_image_transformer_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200_decoder.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "b200_gemm" in syms and "b200_engine_forward_loss" in syms and "b200_engine_generate_greedy" in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (b200_[a-z0-9_]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = sorted(exported - set(header_symbols()))
    assert not extra, f"exported but not declared in the header: {extra}"


def test_loader_binds_every_symbol():
    lib = L.lib()
    assert sorted(L.SIGNATURES) == header_symbols()
    assert lib.b200_version() == 1
    assert lib.b200_last_error() is not None


def test_sass_is_blackwell_native():
    """tcgen05 / TMEM / TMA must be in the shipped SASS (B200_PROFILING.md 'What proves…')."""
    out = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out and "LDTM" in out and "UTMALDG" in out
