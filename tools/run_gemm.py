"""Runs the tcgen05 GEMM at one BASELINE cfg2 shape (for ncu --set full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 12032, 3072, 768            # FFN1 forward of BASELINE configs[1]
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
bias = torch.randn(N, device=dev); D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(6):
    ops.gemm(A, B, bias=bias, act=1, out=D)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gemm(A, B, bias=bias, act=1, out=D)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"gemm M{M} N{N} K{K} bias+relu: {ms*1e3:.1f} us, {2*M*N*K/ms/1e9:.1f} TFLOP/s")
