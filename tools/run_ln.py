"""Times layernorm_bwd / layernorm_fwd / colsum alone at BASELINE cfg2 shapes (M=12032, E=768)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda:0")
M, E = 12032, 768
x = torch.randn(M, E, device=dev).bfloat16(); dy = torch.randn(M, E, device=dev).bfloat16()
g = torch.randn(E, device=dev); mean = torch.randn(M, device=dev); rstd = torch.rand(M, device=dev) + 0.5
dx = torch.empty_like(x); dg = torch.zeros(E, device=dev); db = torch.zeros(E, device=dev); ds = torch.zeros(E, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
t = timeit(lambda: L.check(lib.b200_layernorm_bwd(L.ptr(dy), L.ptr(x), L.ptr(g), L.ptr(mean), L.ptr(rstd), L.ptr(dx), L.ptr(dg), L.ptr(db), L.ptr(ds), M, E, L.cur_stream())))
print(f"layernorm_bwd waves={os.environ.get('B200_LN_BWD_WAVES','4')}: {t:.1f} us  ({3*M*E*2/t/1e3:.0f} GB/s algorithmic)")
