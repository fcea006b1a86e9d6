"""Packed (cu_seqlens) attention against the padded call on the same samples: forward output and dq / dk / dv, for the
caption self attention (packed q and k, causal) and the cross attention (packed q, regular k)."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda:0")

def run(B, H, hd, T, S, self_attn, seed=0, full=False):
    E = H * hd
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(max(1, T // 4), T + 1, (B,), generator=g); lens[0] = T
    if full: lens[:] = T
    cu = torch.zeros(B + 1, dtype=torch.int32); cu[1:] = torch.cumsum(lens, 0).int()
    M = int(cu[-1])
    Tk = T if self_attn else S
    q = torch.randn(B, T, E, generator=g).bfloat16().to(dev); do = torch.randn(B, T, E, generator=g).bfloat16().to(dev)
    k = torch.randn(B, Tk, E, generator=g).bfloat16().to(dev); v = torch.randn(B, Tk, E, generator=g).bfloat16().to(dev)
    tokens = torch.ones(B, T, dtype=torch.int64)
    for b in range(B): tokens[b, int(lens[b]):] = 0
    tokens = tokens.to(dev)
    def call(packed):
        if packed:
            idx = torch.cat([torch.arange(int(lens[b])) + b * T for b in range(B)]).to(dev)
            qq = q.reshape(B * T, E)[idx].contiguous(); dd = do.reshape(B * T, E)[idx].contiguous()
            kk = k.reshape(B * Tk, E)[idx].contiguous() if self_attn else k; vv = v.reshape(B * Tk, E)[idx].contiguous() if self_attn else v
        else:
            qq, dd, kk, vv = q, do, k, v
        o = torch.zeros_like(qq); lse = torch.zeros(B, H, T, device=dev)
        dq, dk, dv = torch.zeros_like(qq), torch.zeros_like(kk), torch.zeros_like(vv)
        a = L.AttnFwdArgs()
        a.q, a.q_bs, a.q_ts = qq.data_ptr(), T * E, E
        a.k, a.k_bs, a.k_ts = kk.data_ptr(), Tk * E, E
        a.v, a.v_bs, a.v_ts = vv.data_ptr(), Tk * E, E
        a.o, a.o_bs, a.o_ts = o.data_ptr(), T * E, E
        a.lse, a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = lse.data_ptr(), B, H, T, Tk, hd, int(self_attn)
        a.key_tokens = tokens.data_ptr() if (self_attn and not packed) else None
        a.pad_idx, a.key_pad_mask, a.scale = 0, None, 1 / math.sqrt(hd)
        cud = cu.to(dev)
        if packed:
            a.cu_q, a.total_q = cud.data_ptr(), M
            if self_attn: a.cu_k, a.total_k = cud.data_ptr(), M
        L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "fwd")
        bw = L.AttnBwdArgs(); bw.f = a
        bw.d_o, bw.do_bs, bw.do_ts = dd.data_ptr(), T * E, E
        bw.dq, bw.dq_bs, bw.dq_ts = dq.data_ptr(), T * E, E
        bw.dk, bw.dk_bs, bw.dk_ts = dk.data_ptr(), Tk * E, E
        bw.dv, bw.dv_bs, bw.dv_ts = dv.data_ptr(), Tk * E, E
        L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "bwd")
        torch.cuda.synchronize()
        if packed:
            def unpack(x, Tn):
                out = torch.zeros(B * Tn, E, device=dev, dtype=x.dtype); out[idx] = x; return out.view(B, Tn, E)
            o, dq = unpack(o, T), unpack(dq, T)
            if self_attn: dk, dv = unpack(dk, Tk), unpack(dv, Tk)
        return o.float(), dq.float(), dk.float(), dv.float()
    # the padded call must not see the PAD rows' dO (the engine's CE gives them zero gradients)
    rowmask = (tokens != 0).unsqueeze(-1)
    do.mul_(rowmask)
    ref = call(False); got = call(True)
    msg = []
    for n, r, x in zip(("o", "dq", "dk", "dv"), ref, got):
        if n in ("o", "dq") or self_attn:
            m = rowmask if n in ("o", "dq") or self_attn else 1
            r = r * m; x = x * m
        err = (r - x).abs().max().item(); msg.append(f"{n} max|diff| {err:.3e} (|ref| {r.abs().max().item():.2f})")
        if err > 1e-2 and n == "o":
            per = (r - x).abs().amax(dim=(1, 2)).tolist(); msg.append("per sample " + " ".join(f"{e:.2f}" for e in per) + f" lens {lens.tolist()}")
            rows = (r - x).abs().amax(dim=2)[1].tolist(); msg.append("sample 1 rows " + " ".join(f"{e:.1f}" for e in rows))
    print(f"B{B} H{H} hd{hd} T{T} S{S} {'self' if self_attn else 'cross'}: " + "  ".join(msg), flush=True)

run(3, 2, 32, 17, 13, False, full=True); run(3, 2, 32, 17, 13, True, full=True)
for hd in (32, 64):
    run(3, 2, hd, 17, 13, True); run(4, 2, hd, 47, 197, True); run(3, 2, hd, 17, 13, False); run(4, 2, hd, 47, 197, False); run(5, 3, hd, 47, 257, False)
