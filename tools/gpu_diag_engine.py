"""GPU diagnostic for the kernels and the engine vs the CPU oracle (run under gpurun).
usage: python tools/gpu_diag_engine.py [ln|attn|fwd|bwd|opt] ..."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200 import _lib as L
from multimodal_image_transformer_b200.engine import DecoderEngine
from oracle import decoder_oracle as O

lib = L.lib()
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def show(name, a, b, tol):
    r = rel(a, b)
    mx = (a.float().cpu() - b.float().cpu()).abs().max().item()
    nan = bool(torch.isnan(a.float()).any())
    print(f"{name:58s} rel_l2 {r:.3e} max_abs {mx:.3e} nan {nan} {'OK' if (r < tol and not nan) else 'FAIL'}", flush=True)
    return r < tol and not nan


def test_ln():
    ok = True
    for rows, E in [(5, 128), (248, 512), (1000, 768), (333, 1024), (64, 2048)]:
        g = torch.Generator().manual_seed(rows)
        x = (torch.randn(rows, E, generator=g) * 2 + 0.5).bfloat16()
        gamma = torch.randn(E, generator=g); beta = torch.randn(E, generator=g)
        dy = torch.randn(rows, E, generator=g).bfloat16()
        xf = x.float().requires_grad_(True); gf = gamma.clone().requires_grad_(True); bf = beta.clone().requires_grad_(True)
        y = torch.nn.functional.layer_norm(xf, (E,), gf, bf, 1e-5)
        y.backward(dy.float())
        xd, dyd = x.to(dev), dy.to(dev)
        yd = torch.empty_like(xd); mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
        L.check(lib.b200_layernorm_fwd(L.ptr(xd), L.ptr(gamma.to(dev)), L.ptr(beta.to(dev)), L.ptr(yd), L.ptr(mean), L.ptr(rstd), rows, E, 1e-5, L.cur_stream()))
        dxd = torch.empty_like(xd); dg = torch.zeros(E, device=dev); db = torch.zeros(E, device=dev); dsum = torch.zeros(E, device=dev)
        gd = gamma.to(dev)
        L.check(lib.b200_layernorm_bwd(L.ptr(dyd), L.ptr(xd), L.ptr(gd), L.ptr(mean), L.ptr(rstd), L.ptr(dxd), L.ptr(dg), L.ptr(db), L.ptr(dsum), rows, E, L.cur_stream()))
        torch.cuda.synchronize()
        ok &= show(f"ln fwd {rows}x{E}", yd, y.detach(), 1e-2)
        ok &= show(f"ln bwd dx {rows}x{E}", dxd, xf.grad, 1e-2)
        ok &= show(f"ln bwd dgamma {rows}x{E}", dg, gf.grad, 1e-3)
        ok &= show(f"ln bwd dbeta {rows}x{E}", db, bf.grad, 1e-3)
    return ok


def attn_ref(q, k, v, causal, keymask):
    # q (B,Tq,H,hd) fp32 etc.
    B, Tq, H, hd = q.shape
    Tk = k.shape[1]
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) / math.sqrt(hd)
    if causal:
        s = s + torch.full((Tq, Tk), float("-inf")).triu(1)
    if keymask is not None:
        s = s.masked_fill(keymask.view(B, 1, 1, Tk), float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v)


def test_attn():
    ok = True
    for (B, H, Tq, Tk, hd, causal) in [(2, 2, 17, 17, 64, 1), (3, 4, 31, 31, 64, 1), (2, 3, 47, 197, 64, 0),
                                       (2, 2, 31, 50, 64, 0), (2, 2, 47, 47, 96, 1), (2, 2, 47, 257, 128, 0),
                                       (1, 2, 99, 99, 128, 1), (2, 2, 20, 1, 64, 0), (2, 2, 33, 40, 32, 0)]:
        g = torch.Generator().manual_seed(B * 100 + Tq + Tk + hd)
        E = H * hd
        q = torch.randn(B, Tq, H, hd, generator=g).bfloat16()
        k = torch.randn(B, Tk, H, hd, generator=g).bfloat16()
        v = torch.randn(B, Tk, H, hd, generator=g).bfloat16()
        do = torch.randn(B, Tq, H, hd, generator=g).bfloat16()
        keymask = torch.zeros(B, Tk, dtype=torch.bool)
        if Tk > 4:
            keymask[0, Tk - 3:] = True
            if causal:
                keymask[-1, 2] = True
        qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
        o_ref = attn_ref(qf, kf, vf, causal, keymask)
        o_ref.backward(do.float())
        qd, kd, vd, dod = (t.to(dev).contiguous() for t in (q, k, v, do))
        od = torch.zeros(B, Tq, E, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(B, H, Tq, device=dev)
        a = L.AttnFwdArgs()
        a.q, a.q_bs, a.q_ts = qd.data_ptr(), Tq * E, E
        a.k, a.k_bs, a.k_ts = kd.data_ptr(), Tk * E, E
        a.v, a.v_bs, a.v_ts = vd.data_ptr(), Tk * E, E
        a.o, a.o_bs, a.o_ts = od.data_ptr(), Tq * E, E
        a.lse = lse.data_ptr()
        a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = B, H, Tq, Tk, hd, causal
        km = keymask.to(torch.uint8).to(dev)
        a.key_tokens, a.pad_idx, a.key_pad_mask = None, 0, km.data_ptr()
        a.scale = 1.0 / math.sqrt(hd)
        L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "attn_fwd")
        torch.cuda.synchronize()
        tag = f"attn B{B} H{H} Tq{Tq} Tk{Tk} hd{hd} c{causal}"
        ok &= show(tag + " fwd", od.view(B, Tq, H, hd), o_ref.detach(), 2e-2)
        bw = L.AttnBwdArgs()
        bw.f = a
        dq = torch.zeros_like(qd); dk = torch.zeros_like(kd); dv = torch.zeros_like(vd)
        bw.d_o, bw.do_bs, bw.do_ts = dod.data_ptr(), Tq * E, E
        bw.dq, bw.dq_bs, bw.dq_ts = dq.data_ptr(), Tq * E, E
        bw.dk, bw.dk_bs, bw.dk_ts = dk.data_ptr(), Tk * E, E
        bw.dv, bw.dv_bs, bw.dv_ts = dv.data_ptr(), Tk * E, E
        L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "attn_bwd")
        torch.cuda.synchronize()
        ok &= show(tag + " dq", dq, qf.grad, 2e-2)
        ok &= show(tag + " dk", dk, kf.grad, 2e-2)
        ok &= show(tag + " dv", dv, vf.grad, 2e-2)
    return ok


CFGS = {
    "tiny": dict(V=1000, E=128, H=2, L=2, F=256, ML=40, B=3, T=17, S=13),
    "cfg1": dict(V=10000, E=512, H=8, L=4, F=2048, ML=100, B=8, T=31, S=50),
    "hd96": dict(V=2000, E=192, H=2, L=2, F=320, ML=64, B=4, T=47, S=197),
}


def make(cfgname, seed=42):
    c = CFGS[cfgname]
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    B, T, S, V, E = c["B"], c["T"], c["S"], c["V"], c["E"]
    tok = torch.randint(4, V, (B, T), generator=g); tok[:, 0] = 1
    tgt = torch.randint(4, V, (B, T), generator=g)
    for b in range(B):
        ln = int(torch.randint(T // 2, T + 1, (1,), generator=g))
        tok[b, ln:] = 0; tgt[b, max(ln - 1, 1):] = 0
    if B > 1:
        tok[1, 3] = 0
    mem = torch.randn(B, S, E, generator=g)
    eng = DecoderEngine(V, E, c["H"], c["L"], c["F"], c["ML"], pad_idx=0, device=dev)
    eng.load(p)
    return c, p, tok, tgt, mem, eng


def test_fwd():
    ok = True
    for name in CFGS:
        c, p, tok, tgt, mem, eng = make(name)
        with torch.no_grad():
            ref = O.decoder_forward(p, tok, mem, None, c["H"])
        got = eng.forward_logits(tok.to(dev), mem.to(dev), None, training=False)
        torch.cuda.synchronize()
        err = (got.cpu() - ref).abs().amax(-1) / ref.abs().amax(-1)
        print(f"{name}: logits max-relative (row-normalised) {err.max().item():.3e}")
        ok &= show(f"{name} logits (inference plan)", got, ref, 2e-2)
        got2 = eng.forward_logits(tok.to(dev), mem.to(dev), None, training=True)
        ok &= show(f"{name} logits (training plan)", got2, ref, 2e-2)
        mpm = torch.zeros(c["B"], c["S"], dtype=torch.bool); mpm[0, c["S"] // 2:] = True
        with torch.no_grad():
            ref = O.decoder_forward(p, tok, mem, mpm, c["H"])
        got = eng.forward_logits(tok.to(dev), mem.to(dev), mpm.to(dev), training=False)
        ok &= show(f"{name} logits + memory padding mask", got, ref, 2e-2)
        with torch.no_grad():
            lref = O.cross_entropy(O.decoder_forward(p, tok, mem, None, c["H"]), tgt, 0)
        lg = eng.forward_loss(tok.to(dev), tgt.to(dev), mem.to(dev), None, 0, training=False).cpu()
        print(f"{name}: loss got {lg[0].item():.6f} ref {lref.item():.6f} rel {abs(lg[0].item()-lref.item())/lref.item():.2e} count {lg[1].item()} ref {(tgt != 0).sum().item()}")
        ok &= abs(lg[0].item() - lref.item()) / lref.item() < 1e-3
    return ok


def test_bwd():
    ok = True
    for name in CFGS:
        c, p, tok, tgt, mem, eng = make(name)
        lref, gref = O.loss_and_grads(p, tok, tgt, mem, None, c["H"])
        eng.zero_grad()
        lg = eng.forward_loss(tok.to(dev), tgt.to(dev), mem.to(dev), None, 0, training=True)
        dmem = eng.backward(want_dmemory=True)
        torch.cuda.synchronize()
        print(f"{name}: loss got {lg[0].item():.6f} ref {lref.item():.6f}")
        worst = 0.0
        for k in gref:
            r = rel(eng.view(k, eng.grads), gref[k])
            worst = max(worst, r)
            if r > 2e-2 or "layers.0.self_attn.in_proj_weight" in k or "token_embedding" in k or "fc_out" in k or "norm1" in k:
                print(f"   grad {k:62s} rel_l2 {r:.3e} {'OK' if r < 3e-2 else 'FAIL'}")
        print(f"{name}: worst grad rel_l2 {worst:.3e}")
        ok &= worst < 3e-2
        # dmemory
        memr = mem.clone().requires_grad_(True)
        O.cross_entropy(O.decoder_forward(p, tok, memr, None, c["H"]), tgt, 0).backward()
        ok &= show(f"{name} dmemory", dmem, memr.grad, 3e-2)
        # autograd-compat path: dlogits given
        eng.zero_grad()
        logits = eng.forward_logits(tok.to(dev), mem.to(dev), None, training=True)
        lgf = logits.detach().clone().requires_grad_(True)
        loss2 = torch.nn.functional.cross_entropy(lgf.view(-1, c["V"]), tgt.to(dev).view(-1), ignore_index=0)
        loss2.backward()
        eng.backward_from_dlogits(lgf.grad)
        torch.cuda.synchronize()
        worst = max(rel(eng.view(k, eng.grads), gref[k]) for k in gref)
        print(f"{name}: compat path worst grad rel_l2 {worst:.3e}")
        ok &= worst < 3e-2
    return ok


def test_opt():
    ok = True
    c, p, tok, tgt, mem, eng = make("tiny")
    state = {}
    pr = {k: v.clone() for k, v in p.items()}
    losses_ref, losses = [], []
    for step in range(5):
        lref, gref = O.loss_and_grads(pr, tok, tgt, mem, None, c["H"])
        O.adamw_step(pr, gref, state, lr=1e-3, max_norm=5.0)
        losses_ref.append(lref.item())
        eng.zero_grad()
        lg = eng.forward_loss(tok.to(dev), tgt.to(dev), mem.to(dev), None, 0, training=True)
        eng.backward()
        eng.adamw_step(lr=1e-3, max_norm=5.0)
        losses.append(lg[0].item())
    torch.cuda.synchronize()
    print("loss trajectory ref", [f"{x:.4f}" for x in losses_ref])
    print("loss trajectory got", [f"{x:.4f}" for x in losses])
    ok &= all(abs(a - b) / b < 5e-3 for a, b in zip(losses, losses_ref))
    worst = max(rel(eng.view(k), pr[k]) for k in pr if k in eng.layout)
    print(f"params after 5 steps: worst rel_l2 {worst:.3e}")
    ok &= worst < 1e-2
    return ok


if __name__ == "__main__":
    which = sys.argv[1:] or ["ln", "attn", "fwd", "bwd", "opt"]
    res = {}
    for w in which:
        res[w] = {"ln": test_ln, "attn": test_attn, "fwd": test_fwd, "bwd": test_bwd, "opt": test_opt}[w]()
    print(res, "ALL OK" if all(res.values()) else "SOME FAILED")
