"""Per-tile time of the tcgen05 GEMM against the number of CTAs it runs on (B200_GEMM_GRID_CAP, one process per cap):
constant per-tile time = a per-SM limit (operand ingest or issue), shrinking per-tile time = a chip-wide limit (L2 slices)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import ops

dev = torch.device("cuda:0")
cap = int(os.environ.get("B200_GEMM_GRID_CAP", "148"))
for (M, N, K) in ((12032, 3072, 768), (12032, 768, 3072), (50432, 1536, 768)):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(A, B, out=D)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(A, B, out=D)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    pairs = min(cap, 148) // 2
    rounds = -(-tiles // pairs)
    print(f"cap {cap:3d} CTAs  M{M} N{N} K{K}: {us:8.1f} us  {2.0*M*N*K/us/1e6:6.0f} TFLOP/s  {tiles} tiles / {pairs} pairs = {rounds} rounds -> {us/rounds:6.2f} us per tile round", flush=True)
