"""GPU diagnostic for the tcgen05 GEMM (run under gpurun): prints error statistics per case.
usage: python tools/gpu_diag_gemm.py [kk|kmn|mnmn|epi|perf ...]"""
import ctypes as C
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda:0")


def run_gemm(A, B, M, N, K, a_mn, b_mn, d_fp32=False, accumulate=False, bias=None, residual=None,
             relu_mask=None, act=0, split_k=1, block_n=0, D=None, check=False):
    if D is None:
        D = torch.zeros(M, N, device=dev, dtype=torch.float32 if d_fp32 else torch.bfloat16)
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_mn_major = A.data_ptr(), A.stride(0), int(a_mn)
    a.B, a.ldb, a.b_mn_major = B.data_ptr(), B.stride(0), int(b_mn)
    a.D, a.ldd, a.d_fp32, a.accumulate = D.data_ptr(), D.stride(0), int(d_fp32), int(accumulate)
    a.bias = bias.data_ptr() if bias is not None else None
    a.residual, a.ldr = (residual.data_ptr(), residual.stride(0)) if residual is not None else (None, 0)
    a.relu_mask, a.ldm = (relu_mask.data_ptr(), relu_mask.stride(0)) if relu_mask is not None else (None, 0)
    a.act, a.split_k, a.block_n = act, split_k, block_n
    fn = lib.b200_gemm_check if check else lib.b200_gemm
    L.check(fn(C.byref(a), L.cur_stream()), "gemm")
    return D


def report(name, got, ref, tol):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-6
    bad = (err > tol * denom).float().mean().item()
    print(f"{name:60s} max_abs_err {err.max().item():.4e} ref_max {denom:.3e} bad_frac {bad:.4f} "
          f"{'OK' if bad == 0 else 'FAIL'}", flush=True)
    if bad > 0:
        idx = (err > tol * denom).nonzero()[:8]
        for i in idx:
            r, c = i.tolist()
            print(f"    [{r},{c}] got {got[r, c].item():.5f} ref {ref[r, c].item():.5f}")
    return bad == 0


def case(name, M, N, K, a_mn, b_mn, **kw):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    Af = torch.randn(M, K, generator=g).to(dev).bfloat16()
    Bf = torch.randn(N, K, generator=g).to(dev).bfloat16()
    ref = Af.float() @ Bf.float().t()
    A = Af.t().contiguous() if a_mn else Af
    B = Bf.t().contiguous() if b_mn else Bf
    ok = True
    for bn in kw.pop("block_ns", (128, 256)):
        if kw.get("d_fp32"):
            D = run_gemm(A, B, M, N, K, a_mn, b_mn, block_n=bn, **kw)
        else:
            D = run_gemm(A, B, M, N, K, a_mn, b_mn, block_n=bn, **kw)
        torch.cuda.synchronize()
        ok &= report(f"{name} M{M} N{N} K{K} bn{bn}", D, ref, 1e-2 if not kw.get("d_fp32") else 2e-3)
    return ok


def main():
    L.check(lib.b200_check_device(0), "check_device")
    which = sys.argv[1:] or ["kk", "kmn", "mnmn", "epi", "perf"]
    ok = True
    if "kk" in which:
        ok &= case("KK", 128, 128, 64, False, False, block_ns=(128,))
        ok &= case("KK", 256, 256, 128, False, False)
        ok &= case("KK", 248, 1000, 96, False, False)
        ok &= case("KK", 1024, 2304, 768, False, False)
        ok &= case("KK fp32", 300, 520, 512, False, False, d_fp32=True)
    if "kmn" in which:
        ok &= case("K-MN (dgrad)", 256, 256, 128, False, True)
        ok &= case("K-MN (dgrad)", 248, 512, 1000, False, True)
    if "mnmn" in which:
        ok &= case("MN-MN (wgrad)", 256, 256, 128, True, True)
        ok &= case("MN-MN (wgrad) fp32", 512, 768, 248, True, True, d_fp32=True)
        ok &= case("MN-MN (wgrad) fp32 splitk4", 768, 768, 4096, True, True, d_fp32=True, accumulate=True, split_k=4)
        ok &= case("MN-MN (wgrad) fp32 auto", 768, 768, 12032, True, True, d_fp32=True, accumulate=True, split_k=0, block_ns=(0,))
        ok &= case("MN-K", 256, 256, 128, True, False)
    if "epi" in which:
        M, N, K = 300, 520, 256
        g = torch.Generator(device="cpu").manual_seed(5)
        A = torch.randn(M, K, generator=g).to(dev).bfloat16()
        B = torch.randn(N, K, generator=g).to(dev).bfloat16()
        bias = torch.randn(N, generator=g).to(dev)
        res = torch.randn(M, N, generator=g).to(dev).bfloat16()
        msk = torch.randn(M, N, generator=g).to(dev).bfloat16()
        base = A.float() @ B.float().t()
        ok &= report("bias", run_gemm(A, B, M, N, K, 0, 0, bias=bias), base + bias, 1e-2)
        ok &= report("bias+relu", run_gemm(A, B, M, N, K, 0, 0, bias=bias, act=1), torch.relu(base + bias), 1e-2)
        ok &= report("bias+gelu", run_gemm(A, B, M, N, K, 0, 0, bias=bias, act=2), torch.nn.functional.gelu(base + bias), 1e-2)
        ok &= report("bias+residual", run_gemm(A, B, M, N, K, 0, 0, bias=bias, residual=res), base + bias + res.float(), 1e-2)
        ok &= report("relu_mask", run_gemm(A, B, M, N, K, 0, 0, relu_mask=msk), base * (msk.float() > 0), 1e-2)
        ok &= report("check kernel", run_gemm(A, B, M, N, K, 0, 0, bias=bias, residual=res, check=True), base + bias + res.float(), 1e-2)
    if "perf" in which:
        for (M, N, K, a_mn, b_mn, f32) in [(12032, 3072, 768, 0, 0, 0), (12032, 768, 3072, 0, 0, 0),
                                           (12032, 2304, 768, 0, 0, 0), (50432, 1536, 768, 0, 0, 0),
                                           (12032, 768, 3072, 0, 1, 0), (3072, 768, 12032, 1, 1, 1),
                                           (768, 768, 12032, 1, 1, 1), (12032, 10000, 768, 0, 0, 0)]:
            A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
            B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
            D = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
            for bn in (128, 256, 0):
                kw = dict(d_fp32=bool(f32), accumulate=bool(f32), split_k=0 if f32 else 1, block_n=bn, D=D)
                for _ in range(3):
                    run_gemm(A, B, M, N, K, a_mn, b_mn, **kw)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    run_gemm(A, B, M, N, K, a_mn, b_mn, **kw)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"perf M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} bn{bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
            # cuBLAS reference
            if not a_mn and not b_mn:
                for _ in range(3): torch.matmul(A, B.t())
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): torch.matmul(A, B.t())
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"   cuBLAS same shape: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    if "small" in which:
        for (M, N, K) in [(512, 768, 768), (512, 2304, 768), (512, 3072, 768), (512, 768, 3072), (512, 10000, 768)]:
            A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
            bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev).bfloat16()
            D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            D32 = torch.zeros(M, N, device=dev, dtype=torch.float32)
            cases = [("bn128", dict(block_n=128, D=D, bias=bias)), ("bn256", dict(block_n=256, D=D, bias=bias)),
                     ("bn128+res", dict(block_n=128, D=D, bias=bias, residual=res))]
            for sk in (2, 3, 6):
                cases.append((f"f32 splitk{sk} bn128", dict(block_n=128, D=D32, d_fp32=True, accumulate=True, split_k=sk)))
                cases.append((f"f32 splitk{sk} bn256", dict(block_n=256, D=D32, d_fp32=True, accumulate=True, split_k=sk)))
            for tag, kw in cases:
                for _ in range(5): run_gemm(A, B, M, N, K, 0, 0, **kw)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50): run_gemm(A, B, M, N, K, 0, 0, **kw)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 50
                print(f"M{M} N{N} K{K} {tag:20s}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    if "noepi" in which:
        for (M, N, K) in [(12032, 3072, 768), (12032, 768, 768), (12032, 768, 3072), (12032, 2304, 768)]:
            A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
            D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev).bfloat16()
            for tag, kw in [("plain", {}), ("bias", dict(bias=bias)), ("bias+res", dict(bias=bias, residual=res)), ("noepi", dict(act=99))]:
                for _ in range(3): run_gemm(A, B, M, N, K, 0, 0, D=D, block_n=256, **kw)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): run_gemm(A, B, M, N, K, 0, 0, D=D, block_n=256, **kw)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"M{M} N{N} K{K} {tag:9s}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    print("ALL OK" if ok else "SOME FAILED")


if __name__ == "__main__":
    main()
