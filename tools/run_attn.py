"""Runs the attention kernels once at BASELINE cfg2 shapes (for ncu captures)."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda:0")
B, H, hd = 256, 12, 64
E = H * hd
def run(Tq, Tk, causal, reps=3):
    q = torch.randn(B, Tq, E, device=dev).bfloat16(); k = torch.randn(B, Tk, E, device=dev).bfloat16()
    v = torch.randn(B, Tk, E, device=dev).bfloat16(); do = torch.randn(B, Tq, E, device=dev).bfloat16()
    o = torch.zeros_like(q); lse = torch.zeros(B, H, Tq, device=dev)
    a = L.AttnFwdArgs()
    a.q, a.q_bs, a.q_ts = q.data_ptr(), Tq * E, E
    a.k, a.k_bs, a.k_ts = k.data_ptr(), Tk * E, E
    a.v, a.v_bs, a.v_ts = v.data_ptr(), Tk * E, E
    a.o, a.o_bs, a.o_ts = o.data_ptr(), Tq * E, E
    a.lse, a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = lse.data_ptr(), B, H, Tq, Tk, hd, causal
    a.key_tokens, a.pad_idx, a.key_pad_mask, a.scale = None, 0, None, 1 / math.sqrt(hd)
    bw = L.AttnBwdArgs(); bw.f = a
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    bw.d_o, bw.do_bs, bw.do_ts = do.data_ptr(), Tq * E, E
    bw.dq, bw.dq_bs, bw.dq_ts = dq.data_ptr(), Tq * E, E
    bw.dk, bw.dk_bs, bw.dk_ts = dk.data_ptr(), Tk * E, E
    bw.dv, bw.dv_bs, bw.dv_ts = dv.data_ptr(), Tk * E, E
    for _ in range(reps):
        L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()))
        L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()))
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream())); e1.record()
    L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream())); e2.record(); torch.cuda.synchronize()
    print(f"Tq{Tq} Tk{Tk} causal{causal}: fwd {e0.elapsed_time(e1)*1e3:.1f} us  bwd {e1.elapsed_time(e2)*1e3:.1f} us")
run(47, 197, 0)
run(47, 47, 1)
