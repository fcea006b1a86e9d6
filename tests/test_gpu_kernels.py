"""Per-kernel parity on the B200, every call through the C ABI.  References are fp32 torch ops on
the SAME bf16-rounded inputs (the kernels' contract: bf16 storage, fp32 accumulate), computed on
the CPU; tolerances are written next to each check."""
import ctypes as C
import math

import pytest
import torch

from multimodal_image_transformer_b200 import _lib as L
from multimodal_image_transformer_b200 import ops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (248, 1000, 96), (1024, 2304, 768), (304, 520, 512), (8, 10000, 512)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("block_n", [128, 256])
def test_gemm_operand_layouts(cuda_dev, M, N, K, a_mn, b_mn, block_n):
    A = _rand((M, K), M + N).bfloat16()
    B = _rand((N, K), K + 1).bfloat16()
    ref = A.float() @ B.float().t()
    Ad = (A.t().contiguous() if a_mn else A).to(cuda_dev)
    Bd = (B.t().contiguous() if b_mn else B).to(cuda_dev)
    out = ops.gemm(Ad, Bd, a_mn=a_mn, b_mn=b_mn, block_n=block_n, M=M, N=N, K=K)
    # bf16 output rounding: 2^-8 relative to the largest magnitude in the row is ample
    assert rel_l2(out, ref) < 4e-3
    out32 = ops.gemm(Ad, Bd, a_mn=a_mn, b_mn=b_mn, block_n=block_n, out_fp32=True, M=M, N=N, K=K)
    assert rel_l2(out32, ref) < 1e-5          # fp32 accumulate of identical bf16 products


def test_gemm_epilogues(cuda_dev):
    M, N, K = 300, 520, 256
    A, B = _rand((M, K), 1).bfloat16(), _rand((N, K), 2).bfloat16()
    bias, res, msk = _rand((N,), 3), _rand((M, N), 4).bfloat16(), _rand((M, N), 5).bfloat16()
    base = A.float() @ B.float().t()
    d = lambda t: t.to(cuda_dev)
    assert rel_l2(ops.gemm(d(A), d(B), bias=d(bias), out_fp32=True), base + bias) < 1e-5
    assert rel_l2(ops.gemm(d(A), d(B), bias=d(bias), act=1, out_fp32=True), torch.relu(base + bias)) < 1e-5
    assert rel_l2(ops.gemm(d(A), d(B), bias=d(bias), act=2, out_fp32=True), torch.nn.functional.gelu(base + bias)) < 1e-4
    # residual / ReLU-mask tiles are fused only into bf16 outputs (bf16 rounding of the result: 2^-8)
    assert rel_l2(ops.gemm(d(A), d(B), bias=d(bias), residual=d(res)), base + bias + res.float()) < 4e-3
    assert rel_l2(ops.gemm(d(A), d(B), relu_mask=d(msk)), base * (msk.float() > 0)) < 4e-3
    assert rel_l2(ops.gemm(d(A), d(B), bias=d(bias), act=1, residual=d(res)), torch.relu(base + bias) + res.float()) < 4e-3
    with pytest.raises(RuntimeError):
        ops.gemm(d(A), d(B), residual=d(res), out_fp32=True)


def test_gemm_split_k_accumulates(cuda_dev):
    M, N, K = 768, 768, 4096                  # wgrad shape: reduction over tokens
    A, B = _rand((M, K), 7).bfloat16(), _rand((N, K), 8).bfloat16()
    ref = A.float() @ B.float().t()
    Ad, Bd = A.t().contiguous().to(cuda_dev), B.t().contiguous().to(cuda_dev)
    for sk in (1, 4, 0):
        out = torch.ones(M, N, device=cuda_dev)
        ops.gemm(Ad, Bd, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=sk, M=M, N=N, K=K)
        assert rel_l2(out - 1.0, ref) < 1e-4


def test_gemm_full_size_against_check_kernel(cuda_dev):
    """BASELINE cfg2 FFN shape: tcgen05 kernel vs the CUDA-core check kernel (a CPU oracle would
    take minutes here) plus linearity: gemm(2A) == 2 gemm(A) exactly in fp32 output."""
    lib = L.lib()
    M, N, K = 12032, 3072, 768
    A = torch.randn(M, K, device=cuda_dev).bfloat16()
    B = torch.randn(N, K, device=cuda_dev).bfloat16()
    bias = torch.randn(N, device=cuda_dev)
    out = ops.gemm(A, B, bias=bias, act=1, out_fp32=True)
    chk = torch.empty_like(out)
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.B, a.ldb = A.data_ptr(), K, B.data_ptr(), K
    a.D, a.ldd, a.d_fp32, a.bias, a.act, a.split_k = chk.data_ptr(), N, 1, bias.data_ptr(), 1, 1
    L.check(lib.b200_gemm_check(C.byref(a), L.cur_stream()))
    torch.cuda.synchronize()
    assert ((out - chk).norm() / chk.norm()).item() < 1e-5
    out2 = ops.gemm((A.float() * 2).bfloat16(), B, out_fp32=True)
    out1 = ops.gemm(A, B, out_fp32=True)
    assert torch.equal(out2, out1 * 2)


def test_gemm_rejects_bad_arguments(cuda_dev):
    A = torch.zeros(64, 64, device=cuda_dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(A, A, N=63)                   # N must be a multiple of 8
    with pytest.raises(RuntimeError):
        ops.gemm(A, A, split_k=4)              # split-K needs an fp32 accumulate output


# ------------------------------------------------------------------ LayerNorm / embedding / reductions
@pytest.mark.parametrize("rows,E", [(5, 64), (248, 512), (1000, 768), (333, 1024), (64, 2048)])
def test_layernorm_fwd_bwd(cuda_dev, rows, E):
    lib = L.lib()
    x = (_rand((rows, E), rows) * 2 + 0.5).bfloat16()
    gamma, beta, dy = _rand((E,), 1), _rand((E,), 2), _rand((rows, E), 3).bfloat16()
    xf, gf, bf = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xf, (E,), gf, bf, 1e-5)
    y.backward(dy.float())
    xd, dyd, gd, bd = x.to(cuda_dev), dy.to(cuda_dev), gamma.to(cuda_dev), beta.to(cuda_dev)
    yd = torch.empty_like(xd)
    mean, rstd = torch.empty(rows, device=cuda_dev), torch.empty(rows, device=cuda_dev)
    L.check(lib.b200_layernorm_fwd(L.ptr(xd), L.ptr(gd), L.ptr(bd), L.ptr(yd), L.ptr(mean), L.ptr(rstd), rows, E, 1e-5, L.cur_stream()))
    dxd, dg, db = torch.empty_like(xd), torch.zeros(E, device=cuda_dev), torch.zeros(E, device=cuda_dev)
    dsum = torch.zeros(E, device=cuda_dev)
    L.check(lib.b200_layernorm_bwd(L.ptr(dyd), L.ptr(xd), L.ptr(gd), L.ptr(mean), L.ptr(rstd), L.ptr(dxd), L.ptr(dg), L.ptr(db), L.ptr(dsum), rows, E, L.cur_stream()))
    torch.cuda.synchronize()
    assert rel_l2(yd, y) < 4e-3 and rel_l2(dxd, xf.grad) < 4e-3          # bf16 output rounding
    assert rel_l2(dg, gf.grad) < 1e-4 and rel_l2(db, bf.grad) < 1e-4   # fp32 outputs
    assert rel_l2(mean, x.float().mean(-1)) < 1e-5
    # fused bias-gradient column sums: fp32 sums of the un-rounded dx; column sums nearly cancel,
    # so the error is measured against the scale of the summands, not of the sum
    scale = xf.grad.abs().sum(0).norm().item()
    assert (dsum.cpu() - xf.grad.sum(0)).norm().item() < 1e-4 * scale


def test_embedding_fwd_bwd(cuda_dev):
    lib = L.lib()
    B, T, E, V = 5, 17, 128, 300
    emb, pe = _rand((V, E), 1, 0.05), _rand((40, E), 2)
    tok = torch.randint(0, V, (B, T), generator=torch.Generator().manual_seed(3))
    tok[0, 5:] = 0
    x = torch.empty(B, T, E, device=cuda_dev, dtype=torch.bfloat16)
    tokd, embd, ped = tok.to(cuda_dev), emb.to(cuda_dev), pe.to(cuda_dev)
    L.check(lib.b200_embed_pe_fwd(L.ptr(tokd), L.ptr(embd), L.ptr(ped), L.ptr(x), B, T, E, V, math.sqrt(E), L.cur_stream()))
    ref = emb[tok] * math.sqrt(E) + pe[:T]
    assert rel_l2(x, ref) < 4e-3
    dx = _rand((B, T, E), 4).bfloat16()
    demb = torch.zeros(V, E, device=cuda_dev)
    dxd = dx.to(cuda_dev)
    L.check(lib.b200_embed_bwd(L.ptr(tokd), L.ptr(dxd), L.ptr(demb), B, T, E, V, 0, math.sqrt(E), L.cur_stream()))
    refg = torch.zeros(V, E).index_add_(0, tok.flatten(), dx.float().view(-1, E) * math.sqrt(E))
    refg[0] = 0
    assert rel_l2(demb, refg) < 1e-5 and float(demb[0].abs().max()) == 0.0


def test_colsum_and_casts(cuda_dev):
    lib = L.lib()
    # ragged against the 256 x 256 block, the 64-column warp strips and the 4-row lanes; a strided view; accumulation
    for M, N, ld in ((1000, 520, 520), (1, 8, 8), (257, 72, 72), (300, 264, 400), (12032, 768, 768)):
        x = _rand((M, ld), 1 + M).bfloat16()
        xd = x.to(cuda_dev)
        out = torch.ones(N, device=cuda_dev)
        L.check(lib.b200_colsum(L.ptr(xd), ld, L.ptr(out), M, N, L.cur_stream()))
        assert rel_l2(out, 1.0 + x[:, :N].float().sum(0)) < 1e-5, (M, N, ld)
    f = _rand((1003,), 2).to(cuda_dev)
    h = torch.empty(1003, device=cuda_dev, dtype=torch.bfloat16)
    L.check(lib.b200_cast_f32_to_bf16(L.ptr(f), L.ptr(h), 1003, L.cur_stream()))
    assert torch.equal(h, f.bfloat16())
    back = torch.empty(1003, device=cuda_dev)
    L.check(lib.b200_cast_bf16_to_f32(L.ptr(h), L.ptr(back), 1003, L.cur_stream()))
    assert torch.equal(back, h.float())


# ------------------------------------------------------------------ attention
def _attn_ref(q, k, v, causal, keymask):
    B, Tq, H, hd = q.shape
    Tk = k.shape[1]
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) / math.sqrt(hd)
    if causal:
        s = s + torch.full((Tq, Tk), float("-inf")).triu(1)
    s = s.masked_fill(keymask.view(B, 1, 1, Tk), float("-inf"))
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v)


@pytest.mark.parametrize("B,H,Tq,Tk,hd,causal", [(2, 2, 17, 17, 64, 1), (3, 4, 31, 31, 64, 1), (2, 3, 47, 197, 64, 0),
                                                 (2, 2, 31, 50, 64, 0), (2, 2, 47, 47, 96, 1), (2, 2, 47, 257, 128, 0),
                                                 (1, 2, 99, 99, 128, 1), (2, 2, 20, 1, 64, 0), (2, 2, 33, 40, 32, 0),
                                                 # tcgen05 kernels (hd 64, Tk >= 65): three key tiles, an exact tile, one
                                                 # query row, a 1-key tail tile, the longest forward, several items per CTA
                                                 (2, 4, 47, 257, 64, 0), (3, 5, 48, 128, 64, 0), (2, 2, 1, 70, 64, 0),
                                                 (5, 7, 33, 129, 64, 0), (2, 4, 47, 288, 64, 0), (32, 12, 47, 197, 64, 0),
                                                 (2, 3, 64, 197, 64, 0)])
def test_attention_fwd_bwd(cuda_dev, B, H, Tq, Tk, hd, causal):
    lib = L.lib()
    E = H * hd
    q, k, v, do = (_rand((B, n, H, hd), i + Tq + Tk).bfloat16() for i, n in enumerate((Tq, Tk, Tk, Tq)))
    keymask = torch.zeros(B, Tk, dtype=torch.bool)
    if Tk > 4:
        keymask[0, Tk - 3:] = True
        if causal:
            keymask[-1, 2] = True
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref = _attn_ref(qf, kf, vf, causal, keymask)
    o_ref.backward(do.float())
    qd, kd, vd, dod = (t.to(cuda_dev).contiguous() for t in (q, k, v, do))
    od = torch.zeros(B, Tq, E, device=cuda_dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, Tq, device=cuda_dev)
    km = keymask.to(torch.uint8).to(cuda_dev)
    a = L.AttnFwdArgs()
    a.q, a.q_bs, a.q_ts = qd.data_ptr(), Tq * E, E
    a.k, a.k_bs, a.k_ts = kd.data_ptr(), Tk * E, E
    a.v, a.v_bs, a.v_ts = vd.data_ptr(), Tk * E, E
    a.o, a.o_bs, a.o_ts = od.data_ptr(), Tq * E, E
    a.lse, a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = lse.data_ptr(), B, H, Tq, Tk, hd, causal
    a.key_tokens, a.pad_idx, a.key_pad_mask, a.scale = None, 0, km.data_ptr(), 1.0 / math.sqrt(hd)
    L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "attn_fwd")
    bw = L.AttnBwdArgs()
    bw.f = a
    dq, dk, dv = torch.zeros_like(qd), torch.zeros_like(kd), torch.zeros_like(vd)
    bw.d_o, bw.do_bs, bw.do_ts = dod.data_ptr(), Tq * E, E
    bw.dq, bw.dq_bs, bw.dq_ts = dq.data_ptr(), Tq * E, E
    bw.dk, bw.dk_bs, bw.dk_ts = dk.data_ptr(), Tk * E, E
    bw.dv, bw.dv_bs, bw.dv_ts = dv.data_ptr(), Tk * E, E
    L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "attn_bwd")
    torch.cuda.synchronize()
    # P and dS are rounded to bf16 before the second product: 2^-8 relative per element
    tol = 6e-3
    assert rel_l2(od.view(B, Tq, H, hd), o_ref) < tol
    for got, ref in ((dq, qf.grad), (dk, kf.grad), (dv, vf.grad)):
        # Tk == 1: dq = dk = 0 exactly; compare absolutely against the gradient scale
        assert (got.float().cpu() - ref).norm().item() < tol * max(ref.norm().item(), 1e-3 * do.float().norm().item())


# ------------------------------------------------------------------ LM head
@pytest.mark.parametrize("M,V,E", [(39, 264, 64), (248, 10000, 512), (1000, 10000, 768)])
def test_lmhead_ce_and_argmax(cuda_dev, M, V, E):
    lib = L.lib()
    x, w, bias = _rand((M, E), 1).bfloat16(), _rand((V, E), 2, 0.05).bfloat16(), _rand((V,), 3, 0.1)
    tg = torch.randint(1, V, (M,), generator=torch.Generator().manual_seed(4))
    tg[::5] = 0
    logits = x.float() @ w.float().t() + bias
    valid = tg != 0
    lse_ref = torch.logsumexp(logits, -1)
    loss_ref = ((lse_ref - logits.gather(1, tg[:, None])[:, 0]) * valid).sum() / valid.sum()
    xd, wd, bd, tgd = x.to(cuda_dev), w.to(cuda_dev), bias.to(cuda_dev), tg.to(cuda_dev)
    n_tiles = (V + 127) // 128
    scratch = torch.empty(3 * M * n_tiles + M, device=cuda_dev)
    row_lse, row_loss = torch.empty(M, device=cuda_dev), torch.empty(M, device=cuda_dev)
    sums = torch.zeros(2, device=cuda_dev)
    L.check(lib.b200_lmhead_ce_fwd(L.ptr(xd), E, L.ptr(wd), E, L.ptr(bd), L.ptr(tgd), M, V, E, 0, L.ptr(row_lse), L.ptr(row_loss),
                                   L.ptr(sums[0:1]), L.ptr(sums[1:2]), L.ptr(scratch), L.cur_stream()), "ce_fwd")
    torch.cuda.synchronize()
    assert sums[1].item() == valid.sum().item()
    assert abs(sums[0].item() / sums[1].item() - loss_ref.item()) < 1e-5 * loss_ref.item()   # fp32 throughout
    assert rel_l2(row_lse, lse_ref) < 1e-6
    inv = torch.full((1,), 1.0 / valid.sum().item(), device=cuda_dev)
    dlog = torch.zeros(M, V, device=cuda_dev, dtype=torch.bfloat16)
    L.check(lib.b200_lmhead_ce_bwd(L.ptr(xd), E, L.ptr(wd), E, L.ptr(bd), L.ptr(tgd), M, V, E, 0, L.ptr(row_lse), L.ptr(inv),
                                   L.ptr(dlog), V, L.cur_stream()), "ce_bwd")
    ref_d = (torch.softmax(logits, -1) - torch.nn.functional.one_hot(tg, V)) * valid[:, None] / valid.sum()
    assert rel_l2(dlog, ref_d) < 4e-3                                       # bf16 output rounding
    ids = torch.empty(M, device=cuda_dev, dtype=torch.int64)
    mx = torch.empty(M, device=cuda_dev)
    L.check(lib.b200_lmhead_argmax(L.ptr(xd), E, L.ptr(wd), E, L.ptr(bd), M, V, E, L.ptr(ids), L.ptr(mx), L.ptr(scratch), L.cur_stream()), "argmax")
    assert torch.equal(ids.cpu(), logits.argmax(-1))
    assert rel_l2(mx, logits.max(-1).values) < 1e-6


def test_argmax_takes_first_index_on_ties(cuda_dev):
    lib = L.lib()
    M, V, E = 4, 520, 64
    x = torch.zeros(M, E, dtype=torch.bfloat16, device=cuda_dev)        # all logits equal the bias
    w = torch.zeros(V, E, dtype=torch.bfloat16, device=cuda_dev)
    bias = torch.zeros(V, device=cuda_dev)
    bias[[300, 129, 400]] = 1.0
    ids = torch.empty(M, device=cuda_dev, dtype=torch.int64)
    scratch = torch.empty(2 * M * 8, device=cuda_dev)
    L.check(lib.b200_lmhead_argmax(L.ptr(x), E, L.ptr(w), E, L.ptr(bias), M, V, E, L.ptr(ids), None, L.ptr(scratch), L.cur_stream()))
    assert ids.tolist() == [129] * M


# ------------------------------------------------------------------ optimizer
def test_adamw_kernel_matches_torch(cuda_dev):
    lib = L.lib()
    n = 100003
    p0, g = _rand((n,), 1), _rand((n,), 2, 3.0)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    pd, gd = p0.to(cuda_dev), g.to(cuda_dev)
    m, v = torch.zeros(n, device=cuda_dev), torch.zeros(n, device=cuda_dev)
    p16 = torch.zeros(n, device=cuda_dev, dtype=torch.bfloat16)
    ss = torch.zeros(1, device=cuda_dev)
    for step in (1, 2, 3):
        ref.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_([ref], 5.0)
        opt.step()
        ss.zero_()
        L.check(lib.b200_grad_sumsq(L.ptr(gd), n, L.ptr(ss), L.cur_stream()))
        L.check(lib.b200_adamw_step(L.ptr(pd), L.ptr(p16), L.ptr(gd), L.ptr(m), L.ptr(v), n, L.ptr(ss), 5.0, 1e-3, 0.9, 0.98, 1e-9,
                                    1e-5, step, L.cur_stream()))
        torch.cuda.synchronize()
        assert abs(math.sqrt(ss.item()) - total.item()) < 1e-4 * total.item()
        assert (pd.cpu() - ref.detach()).abs().max().item() < 2e-6
    assert torch.equal(p16, pd.bfloat16())


def _attn_call(q, k, v, do, H, hd, T, Tk, causal, key_tokens=None, cu_q=None, cu_k=None, total=0, B=None):
    """b200_attn_fwd + b200_attn_bwd through the C ABI; q / k / v / do are [rows, E] bf16 (packed) or [B, T, E]."""
    import ctypes as C
    from multimodal_image_transformer_b200 import _lib as L
    lib = L.lib()
    E = H * hd
    dev = q.device
    o = torch.zeros_like(q)
    lse = torch.zeros(B, H, T, device=dev)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    a = L.AttnFwdArgs()
    a.q, a.q_bs, a.q_ts = q.data_ptr(), T * E, E
    a.k, a.k_bs, a.k_ts = k.data_ptr(), Tk * E, E
    a.v, a.v_bs, a.v_ts = v.data_ptr(), Tk * E, E
    a.o, a.o_bs, a.o_ts = o.data_ptr(), T * E, E
    a.lse, a.B, a.H, a.Tq, a.Tk, a.hd, a.causal = lse.data_ptr(), B, H, T, Tk, hd, int(causal)
    a.key_tokens = key_tokens.data_ptr() if key_tokens is not None else None
    a.pad_idx, a.key_pad_mask, a.scale = 0, None, hd ** -0.5
    if cu_q is not None:
        a.cu_q, a.total_q = cu_q.data_ptr(), total
    if cu_k is not None:
        a.cu_k, a.total_k = cu_k.data_ptr(), total
    L.check(lib.b200_attn_fwd(C.byref(a), L.cur_stream()), "attn_fwd")
    bw = L.AttnBwdArgs()
    bw.f = a
    bw.d_o, bw.do_bs, bw.do_ts = do.data_ptr(), T * E, E
    bw.dq, bw.dq_bs, bw.dq_ts = dq.data_ptr(), T * E, E
    bw.dk, bw.dk_bs, bw.dk_ts = dk.data_ptr(), Tk * E, E
    bw.dv, bw.dv_bs, bw.dv_ts = dv.data_ptr(), Tk * E, E
    L.check(lib.b200_attn_bwd(C.byref(bw), L.cur_stream()), "attn_bwd")
    torch.cuda.synchronize()
    return o, dq, dk, dv


@pytest.mark.parametrize("hd,T,S,self_attn", [(32, 17, 13, True), (32, 47, 197, False), (64, 47, 47, True), (64, 47, 197, False),
                                              (64, 31, 257, False), (96, 47, 197, False)])
def test_packed_attention_is_bit_identical_to_padded(cuda_dev, hd, T, S, self_attn):
    """Packed (cu_seqlens) attention through the C ABI -- every kernel family: mma.sync self / cross, tcgen05 cross forward
    and backward (hd 64, S >= 65) -- returns exactly the padded call's rows for every sample."""
    B, H = 5, 3
    E = H * hd
    Tk = T if self_attn else S
    g = torch.Generator().manual_seed(hd + T)
    lens = torch.randint(max(1, T // 4), T + 1, (B,), generator=g)
    lens[0] = T
    cu = torch.zeros(B + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0).int()
    M = int(cu[-1])
    dev = cuda_dev
    q, do = (torch.randn(B, T, E, generator=g).bfloat16().to(dev) for _ in range(2))
    k, v = (torch.randn(B, Tk, E, generator=g).bfloat16().to(dev) for _ in range(2))
    tokens = torch.ones(B, T, dtype=torch.int64)
    for b in range(B):
        tokens[b, int(lens[b]):] = 0
    tokens = tokens.to(dev)
    rowmask = (tokens != 0).unsqueeze(-1)
    do = do * rowmask                                    # PAD rows carry no gradient (the engine's CE ignores them)
    idx = torch.cat([torch.arange(int(lens[b])) + b * T for b in range(B)]).to(dev)
    cud = cu.to(dev)
    ref = _attn_call(q, k, v, do, H, hd, T, Tk, self_attn, key_tokens=tokens if self_attn else None, B=B)
    qp, dop = q.reshape(B * T, E)[idx].contiguous(), do.reshape(B * T, E)[idx].contiguous()
    kp = k.reshape(B * Tk, E)[idx].contiguous() if self_attn else k
    vp = v.reshape(B * Tk, E)[idx].contiguous() if self_attn else v
    got = _attn_call(qp, kp, vp, dop, H, hd, T, Tk, self_attn, cu_q=cud, cu_k=cud if self_attn else None, total=M, B=B)
    for name, r, x in zip(("o", "dq", "dk", "dv"), ref, got):
        if name in ("o", "dq") or self_attn:
            r = r.reshape(-1, E)[idx]
        assert torch.equal(r.reshape(x.shape), x), name
