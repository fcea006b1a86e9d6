"""Thin per-op wrappers over the C ABI for host code that needs a single kernel (not the engine)."""
import ctypes as C

import torch

from . import _lib as L


def gemm(A, B, *, a_mn=False, b_mn=False, bias=None, residual=None, relu_mask=None, act=0, out=None,
         out_fp32=False, accumulate=False, split_k=1, block_n=0, M=None, N=None, K=None):
    """D[M,N] = epi(A . B^T) on bf16 operands through b200_gemm (see include/b200_decoder.h)."""
    lib = L.lib()
    if M is None:
        M = A.shape[1] if a_mn else A.shape[0]
    if K is None:
        K = A.shape[0] if a_mn else A.shape[1]
    if N is None:
        N = B.shape[1] if b_mn else B.shape[0]
    if out is None:
        out = torch.zeros(M, N, device=A.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_mn_major = A.data_ptr(), A.stride(0), int(a_mn)
    a.B, a.ldb, a.b_mn_major = B.data_ptr(), B.stride(0), int(b_mn)
    a.D, a.ldd, a.d_fp32, a.accumulate = out.data_ptr(), out.stride(0), int(out.dtype == torch.float32), int(accumulate)
    a.bias = bias.data_ptr() if bias is not None else None
    a.residual, a.ldr = (residual.data_ptr(), residual.stride(0)) if residual is not None else (None, 0)
    a.relu_mask, a.ldm = (relu_mask.data_ptr(), relu_mask.stride(0)) if relu_mask is not None else (None, 0)
    a.act, a.split_k, a.block_n = act, split_k, block_n
    L.check(lib.b200_gemm(C.byref(a), L.cur_stream()), "b200_gemm")
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """y = x W^T + b for fp32 host-facing tensors: bf16 operands, fp32 accumulate and output."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]).to(torch.bfloat16).contiguous()
    w = weight.detach().to(torch.bfloat16).contiguous()
    b = bias.detach().float().contiguous()
    y = gemm(x2, w, bias=b, out_fp32=True)
    return y.view(*shp[:-1], weight.shape[0])
