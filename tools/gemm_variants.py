"""Times epilogue / operand-layout variants of the tcgen05 GEMM at the BASELINE configs[1] shapes in one process
(CUDA events, 20 launches each) -- decomposes e.g. the FFN dgrad (MN-major weights + ReLU-mask tile) against the
FFN1 forward of the same M x N x K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def timed(name, fn, flops, reps=20):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:58s} {us:7.1f} us  {flops / us / 1e6:7.0f} TFLOP/s", flush=True)


def run(M, N, K, tag):
    A = torch.randn(M, K, device=dev).bfloat16()
    Bk = torch.randn(N, K, device=dev).bfloat16()          # K-major (forward weights)
    Bn = Bk.t().contiguous()                               # [K, N]: MN-major (dgrad reads the stored weights)
    bias = torch.randn(N, device=dev)
    aux = torch.randn(M, N, device=dev).bfloat16()
    D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    timed(f"{tag} K-major B, no epilogue extras", lambda: ops.gemm(A, Bk, out=D), fl)
    timed(f"{tag} K-major B, bias + relu", lambda: ops.gemm(A, Bk, bias=bias, act=1, out=D), fl)
    timed(f"{tag} K-major B, bias + residual", lambda: ops.gemm(A, Bk, bias=bias, residual=aux, out=D), fl)
    timed(f"{tag} K-major B, relu mask", lambda: ops.gemm(A, Bk, relu_mask=aux, out=D), fl)
    timed(f"{tag} MN-major B, no epilogue extras", lambda: ops.gemm(A, Bn, b_mn=True, out=D), fl)
    timed(f"{tag} MN-major B, residual", lambda: ops.gemm(A, Bn, b_mn=True, residual=aux, out=D), fl)
    timed(f"{tag} MN-major B, relu mask", lambda: ops.gemm(A, Bn, b_mn=True, relu_mask=aux, out=D), fl)


print("B200_GEMM_EPI_GROUPS =", os.environ.get("B200_GEMM_EPI_GROUPS", "(default)"))
run(12032, 3072, 768, "M12032 N3072 K768 ")
run(12032, 768, 768, "M12032 N768  K768 ")
run(12032, 768, 3072, "M12032 N768  K3072")
run(12032, 2304, 768, "M12032 N2304 K768 ")
