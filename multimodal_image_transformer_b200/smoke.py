"""One small invocation of the hot path on cuda:0 (train step + greedy decode), checked against the
CPU oracle.  Called by __graft_entry__.smoke(); the oracle import is test infrastructure (allowed
here and nowhere else in the package)."""
import os
import sys

import torch


def run() -> None:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import decoder_oracle as O
    from .engine import DecoderEngine

    if not torch.cuda.is_available():
        raise RuntimeError("smoke(): no CUDA device; the B200 path has no CPU fallback")
    dev = torch.device("cuda", 0)
    # head dim 64 and 77 memory tokens: the cross attention takes the tcgen05 / TMEM kernels (csrc/attention_tc.cu), the
    # caption's self attention the mma.sync ones, every Linear the tcgen05 GEMM
    V, E, H, Ln, F, ML, B, T, S = 1000, 128, 2, 2, 256, 40, 4, 17, 77
    p = O.init_params(V, E, H, Ln, F, ML, seed=42)
    g = torch.Generator().manual_seed(1)
    tok = torch.randint(4, V, (B, T), generator=g)
    tok[:, 0] = 1
    tok[0, 11:] = 0
    tgt = torch.randint(4, V, (B, T), generator=g)
    tgt[0, 10:] = 0
    mem = torch.randn(B, S, E, generator=g)

    eng = DecoderEngine(V, E, H, Ln, F, ML, pad_idx=0, device=dev)
    eng.load(p)
    with torch.no_grad():
        ref_logits = O.decoder_forward(p, tok, mem, None, H)
    logits = eng.forward_logits(tok.to(dev), mem.to(dev), None).cpu()
    err = ((logits - ref_logits).abs().amax(-1) / ref_logits.abs().amax(-1)).max().item()
    assert err < 2e-2, f"logits max-relative error {err}"

    lref, gref = O.loss_and_grads(p, tok, tgt, mem, None, H)
    eng.zero_grad()
    out = eng.forward_loss(tok.to(dev), tgt.to(dev), mem.to(dev), None, 0, training=True)
    eng.backward()
    loss = out[0].item()
    assert abs(loss - lref.item()) < 1e-3 * lref.item(), (loss, lref.item())
    # the packed (var-len) path on the same batch and weights: same loss
    lens = DecoderEngine.packed_lengths(tok, 0)
    out_pk = eng.forward_loss(tok.to(dev), tgt.to(dev), mem.to(dev), None, 0, training=False, lengths=lens)
    assert abs(out_pk[0].item() - loss) < 1e-5 * abs(loss), (out_pk[0].item(), loss)
    eng.adamw_step(lr=1e-3, max_norm=5.0)
    gw = eng.view("fc_out.weight", eng.grads).cpu()
    gerr = ((gw - gref["fc_out.weight"]).norm() / gref["fc_out.weight"].norm()).item()
    assert gerr < 2e-2, f"fc_out.weight gradient error {gerr}"

    eng.load(p)
    # generation on a 13-token memory (with random-init weights longer memories put the first argmax within bf16 rounding
    # of a tie; the near-tie rule lives in the tests, the smoke check wants exact ids)
    mem_d = mem[:2, :13].contiguous()
    with torch.no_grad():
        ref_ids = O.greedy_generate(p, mem_d, 1, 2, 8, H)
    eng.decode_begin(mem_d.to(dev), None, beam=1, max_len=8)
    toks, lens = eng.generate_greedy(1, 2, 8, 0)
    got = [toks[b, :int(lens[b])].tolist() for b in range(2)]
    assert got == ref_ids, (got, ref_ids)
    torch.cuda.synchronize()
    print(f"smoke ok: logits err {err:.2e}, loss {loss:.4f} (oracle {lref.item():.4f}), grad err {gerr:.2e}, greedy ids match")
