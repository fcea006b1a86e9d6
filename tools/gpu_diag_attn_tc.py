"""Bring-up / regression diagnostic of the tcgen05 attention kernels (csrc/attention_tc.cu) through the C ABI:
forward (O, log-sum-exp) and backward (dQ, dK, dV) against plain fp32 torch math on the same bf16-rounded inputs, for
the shapes of the train step (cross attention 47 x 197 / 47 x 257 / 31 x 50, causal self attention with key padding,
a one-token memory), plus per-launch timings at the BASELINE cfg2 / cfg5 sizes.  Prints one line per case."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from multimodal_image_transformer_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = torch.device("cuda:0")
hd = 64


def ref_attn(q, k, v, H, causal, key_mask):
    B, Tq, E = q.shape
    Tk = k.shape[1]
    qh = q.float().view(B, Tq, H, hd).transpose(1, 2)
    kh = k.float().view(B, Tk, H, hd).transpose(1, 2)
    vh = v.float().view(B, Tk, H, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Tq, Tk, dtype=torch.bool, device=q.device), 1), float("-inf"))
    if key_mask is not None:
        s = s.masked_fill(key_mask.view(B, 1, 1, Tk), float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.softmax(s, -1)
    o = (p @ vh).transpose(1, 2).reshape(B, Tq, E)
    return o, lse


def case(B, H, Tq, Tk, causal, mask, bwd=True, reps=0, seed=0):
    E = H * hd
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(B, Tq, E, generator=g).to(dev, torch.bfloat16)
    k = torch.randn(B, Tk, E, generator=g).to(dev, torch.bfloat16)
    v = torch.randn(B, Tk, E, generator=g).to(dev, torch.bfloat16)
    do = torch.randn(B, Tq, E, generator=g).to(dev, torch.bfloat16)
    key_mask, tokens, pad_mask = None, None, None
    if mask == "tokens":
        tokens = torch.randint(4, 100, (B, Tk), generator=g)
        for b in range(B):
            ln = int(torch.randint(max(1, Tk // 2), Tk + 1, (1,), generator=g))
            tokens[b, ln:] = 0
        if B > 1 and Tk > 3:
            tokens[1, 3] = 0
        tokens = tokens.to(dev)
        key_mask = tokens == 0
    elif mask == "pad":
        pad_mask = torch.zeros(B, Tk, dtype=torch.uint8)
        pad_mask[0, Tk // 2:] = 1
        pad_mask = pad_mask.to(dev)
        key_mask = pad_mask.bool()
    o = torch.full_like(q, float("nan"))
    lse = torch.full((B, H, Tq), float("nan"), device=dev)
    a = L.AttnFwdArgs()
    a.q, a.q_bs, a.q_ts = q.data_ptr(), Tq * E, E
    a.k, a.k_bs, a.k_ts = k.data_ptr(), Tk * E, E
    a.v, a.v_bs, a.v_ts = v.data_ptr(), Tk * E, E
    a.o, a.o_bs, a.o_ts = o.data_ptr(), Tq * E, E
    a.lse, a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = lse.data_ptr(), B, H, Tq, Tk, hd, causal
    a.key_tokens = tokens.data_ptr() if tokens is not None else None
    a.pad_idx = 0
    a.key_pad_mask = pad_mask.data_ptr() if pad_mask is not None else None
    a.scale = 1 / math.sqrt(hd)
    L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "attn_fwd")
    torch.cuda.synchronize()
    qr = q.clone().float().requires_grad_(True)
    kr = k.clone().float().requires_grad_(True)
    vr = v.clone().float().requires_grad_(True)
    o_ref, lse_ref = ref_attn(qr, kr, vr, H, causal, key_mask)
    finite = torch.isfinite(lse_ref)
    eo = (o.float() - o_ref).abs().max().item()
    el = (lse[finite] - lse_ref[finite]).abs().max().item() if finite.any() else 0.0
    nan_o = int(torch.isnan(o.float()).sum())
    msg = f"B{B} H{H} Tq{Tq} Tk{Tk} causal{causal} mask={mask}: fwd max|dO| {eo:.3e} max|dlse| {el:.3e} nan {nan_o}"
    ok = eo < 2e-2 and el < 1e-3 and nan_o == 0
    if bwd:
        dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
        bw = L.AttnBwdArgs()
        bw.f = a
        bw.d_o, bw.do_bs, bw.do_ts = do.data_ptr(), Tq * E, E
        bw.dq, bw.dq_bs, bw.dq_ts = dq.data_ptr(), Tq * E, E
        bw.dk, bw.dk_bs, bw.dk_ts = dk.data_ptr(), Tk * E, E
        bw.dv, bw.dv_bs, bw.dv_ts = dv.data_ptr(), Tk * E, E
        L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "attn_bwd")
        torch.cuda.synchronize()
        o_ref.backward(do.float())
        errs = []
        for name, got, ref in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
            # relative to the gradient's own norm, with a floor (a one-key softmax has dS = 0 exactly: dq = dk = 0)
            rel = ((got.float() - ref).norm() / ref.norm().clamp_min(1e-3 * do.float().norm())).item()
            errs.append(rel)
            msg += f" {name} rel {rel:.3e}"
            ok = ok and rel < 1e-2 and not torch.isnan(got.float()).any()
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    if reps:
        for fn, nm in ((lambda: L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "f"), "fwd"),) + \
                      (((lambda: L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "b"), "bwd"),) if bwd else ()):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"     {nm}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch", flush=True)
    return ok


def trace(B, H, Tq, Tk, causal):
    """per-item pipeline timeline of CTA 0 (cycles relative to the first producer issue)"""
    E = H * hd
    q = torch.randn(B, Tq, E, device=dev).bfloat16()
    k = torch.randn(B, Tk, E, device=dev).bfloat16()
    v = torch.randn(B, Tk, E, device=dev).bfloat16()
    o = torch.zeros_like(q)
    lse = torch.zeros(B, H, Tq, device=dev)
    a = L.AttnFwdArgs()
    a.q, a.q_bs, a.q_ts = q.data_ptr(), Tq * E, E
    a.k, a.k_bs, a.k_ts = k.data_ptr(), Tk * E, E
    a.v, a.v_bs, a.v_ts = v.data_ptr(), Tk * E, E
    a.o, a.o_bs, a.o_ts = o.data_ptr(), Tq * E, E
    a.lse, a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = lse.data_ptr(), B, H, Tq, Tk, hd, causal
    a.key_tokens, a.pad_idx, a.key_pad_mask, a.scale = None, 0, None, 1 / math.sqrt(hd)
    for _ in range(3):
        L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "attn_fwd")
    buf = torch.zeros(32, 16, dtype=torch.int64, device=dev)
    L.check(lib.b200_attn_tc_trace(buf.data_ptr()), "trace")
    L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "attn_fwd")
    torch.cuda.synchronize()
    L.check(lib.b200_attn_tc_trace(None), "trace")
    t = buf.cpu()
    t0 = int(t[0, 0])
    print(f"trace Tq{Tq} Tk{Tk} causal{causal}: item: load_issue S_issued P_seen PV_issued | s_full pass1 pass2 o_full epi_done")
    for i in range(min(22, B * H // 148 + 1)):
        r = [int(x) - t0 if int(x) else -1 for x in t[i, :13]]
        print(f"  {i:2d}: {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} | {r[4]:7d} {r[5]:7d} {r[6]:7d} {r[7]:7d} {r[8]:7d} | P published by quarter 0-3: {r[9]:7d} {r[10]:7d} {r[11]:7d} {r[12]:7d}")


def trace_bwd(B, H, Tq, Tk, causal):
    """per-unit (128-key tile) pipeline timeline of CTA 0 of the backward kernel"""
    E = H * hd
    q, do = (torch.randn(B, Tq, E, device=dev).bfloat16() for _ in range(2))
    k, v = (torch.randn(B, Tk, E, device=dev).bfloat16() for _ in range(2))
    o = torch.zeros_like(q)
    lse = torch.zeros(B, H, Tq, device=dev)
    a = L.AttnFwdArgs()
    a.q, a.q_bs, a.q_ts = q.data_ptr(), Tq * E, E
    a.k, a.k_bs, a.k_ts = k.data_ptr(), Tk * E, E
    a.v, a.v_bs, a.v_ts = v.data_ptr(), Tk * E, E
    a.o, a.o_bs, a.o_ts = o.data_ptr(), Tq * E, E
    a.lse, a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = lse.data_ptr(), B, H, Tq, Tk, hd, causal
    a.key_tokens, a.pad_idx, a.key_pad_mask, a.scale = None, 0, None, 1 / math.sqrt(hd)
    L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "attn_fwd")
    dq, dk, dv = (torch.zeros_like(t) for t in (q, k, v))
    bw = L.AttnBwdArgs()
    bw.f = a
    bw.d_o, bw.do_bs, bw.do_ts = do.data_ptr(), Tq * E, E
    bw.dq, bw.dq_bs, bw.dq_ts = dq.data_ptr(), Tq * E, E
    bw.dk, bw.dk_bs, bw.dk_ts = dk.data_ptr(), Tk * E, E
    bw.dv, bw.dv_bs, bw.dv_ts = dv.data_ptr(), Tk * E, E
    for _ in range(3):
        L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "attn_bwd")
    buf = torch.zeros(32, 16, dtype=torch.int64, device=dev)
    L.check(lib.b200_attn_tc_trace(buf.data_ptr()), "trace")
    L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "attn_bwd")
    torch.cuda.synchronize()
    L.check(lib.b200_attn_tc_trace(None), "trace")
    t = buf.cpu()
    t0 = int(t[0, 0])
    print(f"bwd trace Tq{Tq} Tk{Tk} causal{causal}: unit: load_issue S_issued | st_full_seen pds_free published | out_ready out_issued | out_full_seen stored | stats")
    for u in range(32):
        r = [int(x) - t0 if int(x) else -1 for x in t[u, :16]]
        print(f"  {u:2d}: {r[0]:7d} {r[1]:7d} | {r[2]:7d} {r[3]:7d} {r[4]:7d} | {r[5]:7d} {r[6]:7d} | {r[7]:7d} {r[8]:7d} | {r[9]:7d}"
              f" || V load issue {r[14]:7d}; S issuer: at unit {r[10]:7d} K seen {r[11]:7d} V seen {r[12]:7d} stage free {r[13]:7d}")


def main():
    bwd = "--no-bwd" not in sys.argv
    if "--trace-bwd" in sys.argv:
        trace_bwd(256, 12, 47, 197, 0)
        trace_bwd(256, 12, 47, 47, 1)
        return
    if "--bench" in sys.argv:          # one big case only (ncu captures): --bench cross|self|cfg5
        which = sys.argv[sys.argv.index("--bench") + 1]
        shp = {"cross": (256, 12, 47, 197, 0, None), "self": (256, 12, 47, 47, 1, "tokens"), "cfg5": (64, 16, 47, 257, 0, None)}[which]
        case(*shp, bwd, reps=5)
        return
    if "--trace" in sys.argv:
        trace(256, 12, 47, 197, 0)
        trace(256, 12, 47, 47, 1)
        return
    ok = True
    ok &= case(2, 2, 17, 13, 0, None, bwd)
    ok &= case(3, 2, 47, 197, 0, None, bwd)
    ok &= case(3, 2, 47, 197, 0, "pad", bwd)
    ok &= case(2, 4, 47, 257, 0, None, bwd)
    ok &= case(8, 8, 31, 50, 0, None, bwd)
    ok &= case(4, 2, 47, 47, 1, "tokens", bwd)
    ok &= case(3, 2, 17, 17, 1, "tokens", bwd)
    ok &= case(2, 2, 31, 1, 0, None, bwd)
    ok &= case(5, 3, 64, 64, 1, "tokens", bwd)
    ok &= case(256, 12, 47, 197, 0, None, bwd, reps=20)
    ok &= case(256, 12, 47, 47, 1, "tokens", bwd, reps=20)
    ok &= case(64, 16, 47, 257, 0, None, bwd, reps=20)
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
