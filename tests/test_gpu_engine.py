"""Whole-path parity on the B200: the CUDA engine (through the C ABI) and the drop-in modules
against the CPU oracle and the committed golden vectors of the unmodified reference.

Tolerances (BASELINE.json north_star / SURVEY 8c):
  logits  <= 2e-2 max-relative per row (bf16 storage, fp32 accumulate)
  loss    <= 1e-3 relative
  grads   per tensor: relative L2 <= 1e-1 and cosine >= 0.995 against BOTH the fp32 oracle and
          the bf16-emulating oracle.  Why not tighter: a ReLU unit whose pre-activation is within
          bf16 rounding of zero flips its mask and perturbs the gradient by O(sqrt(flip fraction)).
          Measured on the B200 (profiles/r01_grad_noise.log): engine-vs-fp32, engine-vs-emulated and
          emulated-vs-fp32 all sit at 2-7e-2 (cosine 0.998-0.9999), i.e. two bf16 evaluations of the
          reference differ from each other as much as the engine differs from either; tensors next
          to the loss (fc_out, last layer's linear2/norm3), which see no mask flips, agree to <= 7e-3.
          The tight (4e-3) checks of the backward MATH are the per-kernel tests in
          test_gpu_kernels.py (attention / LayerNorm / GEMM dgrad+wgrad / CE on identical inputs).
  greedy  identical token ids on >= 99 % of samples
"""
import pytest
import torch

from oracle import decoder_oracle as O
from tests.helpers import CFGS, golden_params, load_golden, make_engine, rel_l2, row_max_rel, synth

pytestmark = pytest.mark.gpu


def _grad_close(got, ref, rel=1e-1, cos=0.995):
    got = got.detach().float().cpu().flatten()
    ref = ref.detach().float().cpu().flatten()
    c = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    return rel_l2(got, ref) < rel and c > cos


def _case(name, seed=42):
    c = CFGS[name]
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=seed)
    return c, p, synth(c, seed + 1)


@pytest.mark.parametrize("name", ["nano", "tiny", "hd96", "cfg1", "cfg2s", "cfg5s"])
def test_forward_logits_and_loss_vs_oracle(cuda_dev, name):
    c, p, (tok, tgt, mem, mpm) = _case(name)
    eng = make_engine(c, p, cuda_dev)
    with torch.no_grad():
        ref = O.decoder_forward(p, tok, mem, None, c["H"])
        ref_m = O.decoder_forward(p, tok, mem, mpm, c["H"])
    got = eng.forward_logits(tok.to(cuda_dev), mem.to(cuda_dev), None, training=False)
    # 2e-2 (north_star) up to 6 layers.  The 12-layer cfg5 stack sits AT the bf16 noise floor: over four seeds the
    # bf16-emulating oracle itself is 1.4e-2 .. 2.07e-2 away from the fp32 oracle and the engine 1.6e-2 .. 2.08e-2
    # (profiles/r02_logit_noise.txt), so that stack gets 2.5e-2 and must also stay within 1.25x of the emulated floor.
    tol = 2e-2 if c["L"] <= 6 else 2.5e-2
    assert row_max_rel(got, ref) < tol
    if c["L"] > 6:
        with torch.no_grad():
            emu = O.decoder_forward(p, tok, mem, None, c["H"], emulate_bf16=True)
        assert row_max_rel(got, ref) < 1.25 * max(row_max_rel(emu, ref), 1.6e-2)
    got_t = eng.forward_logits(tok.to(cuda_dev), mem.to(cuda_dev), None, training=True)
    assert torch.equal(got, got_t)                     # same kernels, different workspace plan
    got_m = eng.forward_logits(tok.to(cuda_dev), mem.to(cuda_dev), mpm.to(cuda_dev), training=False)
    assert row_max_rel(got_m, ref_m) < tol
    lref = O.cross_entropy(ref, tgt, 0).item()
    lg = eng.forward_loss(tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev), None, 0, training=False).cpu()
    assert abs(lg[0].item() - lref) < 1e-3 * lref
    assert lg[1].item() == (tgt != 0).sum().item()


@pytest.mark.parametrize("name", ["nano", "cfg1", "cfg2"])
def test_forward_vs_reference_golden(cuda_dev, name):
    g = load_golden(name)
    c = g["config"]
    eng = make_engine(c, golden_params(g), cuda_dev)
    tok, mem = g["tokens"].to(cuda_dev), g["memory"].to(cuda_dev)
    got = eng.forward_logits(tok, mem, None).cpu()
    got_m = eng.forward_logits(tok, mem, g["mem_pad"].to(cuda_dev)).cpu()
    if c["full"]:
        assert row_max_rel(got, g["logits"]) < 2e-2 and row_max_rel(got_m, g["logits_mem_pad"]) < 2e-2
    else:
        err = (got[:, :, ::97] - g["logits_sub"]).abs().amax(-1) / g["logits_rowmax"]
        assert err.max().item() < 2e-2
    loss = eng.forward_loss(tok, g["targets"].to(cuda_dev), mem, None, 0).cpu()[0].item()
    assert abs(loss - g["loss"]) < 1e-3 * g["loss"]


@pytest.mark.parametrize("name", ["nano", "tiny", "hd96", "cfg1", "cfg5s"])
def test_gradients_vs_oracle(cuda_dev, name):
    c, p, (tok, tgt, mem, _) = _case(name)
    eng = make_engine(c, p, cuda_dev)
    lref, g32 = O.loss_and_grads(p, tok, tgt, mem, None, c["H"])
    _, g16 = O.loss_and_grads(p, tok, tgt, mem, None, c["H"], emulate_bf16=True)
    eng.zero_grad()
    out = eng.forward_loss(tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev), None, 0, training=True)
    dmem = eng.backward(want_dmemory=True)
    torch.cuda.synchronize()
    assert abs(out[0].item() - lref.item()) < 1e-3 * lref.item()
    # bf16 storage noise grows with depth: the 12-layer cfg5 stack sits at 1.0e-1 against the fp32 oracle for the
    # deepest tensor (the embedding) while staying inside 1e-1 of the bf16-emulating oracle
    rel32 = 1.5e-1 if c["L"] > 6 else 1e-1
    for k in g32:
        got = eng.view(k, eng.grads)
        assert _grad_close(got, g16[k]), (k, rel_l2(got, g16[k]))
        assert _grad_close(got, g32[k], rel=rel32, cos=0.99 if c["L"] > 6 else 0.995), (k, rel_l2(got, g32[k]))
        if k.startswith("fc_out"):
            assert rel_l2(got, g32[k]) < 1.5e-2, (k, rel_l2(got, g32[k]))   # no ReLU between it and the loss
    assert float(eng.view("token_embedding.weight", eng.grads)[0].abs().max()) == 0.0
    memr = mem.clone().requires_grad_(True)
    O.cross_entropy(O.decoder_forward(p, tok, memr, None, c["H"], emulate_bf16=True), tgt, 0).backward()
    assert _grad_close(dmem, memr.grad)
    # backward accumulates (optimizer.zero_grad() semantics of train.py:80): a second pass doubles
    first = eng.grads.clone()
    eng.forward_loss(tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev), None, 0, training=True)
    eng.backward()
    assert rel_l2(eng.grads, 2 * first) < 1e-3


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_gradients_vs_reference_golden_subsampled(cuda_dev, name):
    """Gradients at the benchmark decoder shapes (cfg2 = BASELINE configs[1] at batch 4) against the fixtures
    written by the UNMODIFIED reference (tests/golden/make_golden.py: `grads_sub` = 512 evenly spaced entries
    of every parameter gradient, `grad_norm` = its L2 norm).  Tolerances as in test_gradients_vs_oracle, except that
    the rel-L2 of a 512-entry SUBSAMPLE is a noisy estimate of the whole-tensor figure (which sits at 5-7e-2 for the
    ReLU-adjacent tensors, profiles/r01_grad_noise.log): bound 1.5e-1 on the subsample, 5e-2 on the tensor norm."""
    g = load_golden(name)
    c = g["config"]
    eng = make_engine(c, golden_params(g), cuda_dev)
    tok, tgt, mem = (g[k].to(cuda_dev) for k in ("tokens", "targets", "memory"))
    eng.zero_grad()
    out = eng.forward_loss(tok, tgt, mem, None, 0, training=True)
    eng.backward()
    torch.cuda.synchronize()
    assert abs(out[0].item() - g["loss"]) < 1e-3 * g["loss"]
    worst = {}
    for k, ref_sub in g["grads_sub"].items():
        got = eng.view(k, eng.grads).detach().cpu()
        sub = got.flatten()[::max(1, got.numel() // 512)][:512]
        ref_norm = g["grad_norm"][k]
        if ref_norm < 1e-6 * max(g["grad_norm"].values()):      # key-bias gradients: exactly zero in exact arithmetic
            assert float(got.norm()) <= 1e-3 * max(g["grad_norm"].values()), k
            continue
        assert abs(float(got.norm()) - ref_norm) < 5e-2 * ref_norm, (k, float(got.norm()), ref_norm)
        expect_sub = ref_norm * (sub.numel() / got.numel()) ** 0.5          # norm of a typical 512-entry subsample
        if float(ref_sub.norm()) < 0.05 * expect_sub:      # the sampled entries are (near) zero, e.g. a column of a dead ReLU unit
            assert float(sub.norm()) < 0.1 * expect_sub, (k, float(sub.norm()), expect_sub)
            continue
        assert _grad_close(sub, ref_sub, rel=1.5e-1, cos=0.99), (k, rel_l2(sub, ref_sub))
        if k.startswith("fc_out"):
            assert rel_l2(sub, ref_sub) < 1.5e-2, (k, rel_l2(sub, ref_sub))   # no ReLU between it and the loss
        worst[k] = rel_l2(sub, ref_sub)
    assert len(worst) > 10


def test_dp_bucket_path_equals_full_batch_backward(cuda_dev):
    """The data-parallel code path on one GPU (dp.DataParallel.backward_and_allreduce without the collective):
    two half-batches with UNEQUAL non-PAD counts, each run as forward_loss -> backward_parts(i, i, 1/global count)
    bucket by bucket, summed, must equal backward() of the concatenated batch (CrossEntropyLoss is a mean over ALL
    non-PAD targets, reference train.py:327,90)."""
    c = dict(CFGS["cfg1"], B=8)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=42)
    tok, tgt, mem, _ = synth(c, 43)
    tgt[:3, 9:] = 0                                    # rank 0's half carries far fewer targets
    eng = make_engine(c, p, cuda_dev)
    tokd, tgtd, memd = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    eng.zero_grad()
    full = eng.forward_loss(tokd, tgtd, memd, None, 0, training=True).clone()
    eng.backward()
    g_full = eng.grads.clone()
    n_buckets = len(eng.grad_buckets())
    assert n_buckets == c["L"] + 2
    halves = [(0, 4), (4, 8)]
    counts = [(tgt[a:b] != 0).sum().item() for a, b in halves]
    assert counts[0] != counts[1]
    inv = torch.tensor([1.0 / sum(counts)], device=cuda_dev)
    g_sum = torch.zeros_like(g_full)
    loss_sum = 0.0
    for (a, b), n in zip(halves, counts):
        eng.zero_grad()
        out = eng.forward_loss(tokd[a:b], tgtd[a:b], memd[a:b], None, 0, training=True)
        assert out[1].item() == n
        loss_sum += out[0].item() * n
        for i in range(n_buckets):
            eng.backward_parts(i, i, inv)
        g_sum += eng.grads
    assert abs(loss_sum / sum(counts) - full[0].item()) < 1e-5 * full[0].item()
    assert rel_l2(g_sum, g_full) < 2e-3, rel_l2(g_sum, g_full)       # bf16 dlogits rounding + fp32 summation order
    for off, cnt in eng.grad_buckets():
        assert rel_l2(g_sum[off:off + cnt], g_full[off:off + cnt]) < 5e-3


_LONG = dict(V=264, E=128, H=2, L=2, F=256, ML=100, B=2, T=99, S=197)   # reference MAX_SEQ_LEN: two-phase attention backward


_TCDROP = dict(V=1000, E=128, H=2, L=2, F=256, ML=100, B=3, T=47, S=197)   # head dim 64, S >= 65: the tcgen05 cross attention


@pytest.mark.parametrize("name", ["tiny", "hd96", "cfg1", "long", "tcdrop"])
def test_dropout_matches_oracle_with_same_masks(cuda_dev, name):
    """Dropout p=0.1 at the reference's 1 + 6 L sites (decoder.py:72; transformer.py:1175,1195,1199;
    attention probabilities, functional.py:6682).  torch's random stream cannot be reproduced, so
    the oracle applies torch's dropout SEMANTICS with the CUDA path's counter-based masks
    (oracle.DropSpec) and loss / logits / every gradient are compared as in the p=0 tests."""
    c = _LONG if name == "long" else (_TCDROP if name == "tcdrop" else CFGS[name])
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=42)
    tok, tgt, mem, mpm = synth(c, 43)
    eng = make_engine(c, p, cuda_dev)
    eng.set_dropout(0.1, seed=77)
    tokd, tgtd, memd = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    # eval-mode forwards ignore dropout entirely
    with torch.no_grad():
        ref0 = O.decoder_forward(p, tok, mem, None, c["H"])
    assert row_max_rel(eng.forward_logits(tokd, memd, None, training=False), ref0) < 2e-2
    assert eng.dropout_state() == (77, 0)
    # training forward: logits under mask set #1
    got = eng.forward_logits(tokd, memd, mpm.to(cuda_dev), training=True)
    assert eng.dropout_state() == (77, 1)
    with torch.no_grad():
        ref = O.decoder_forward(p, tok, mem, mpm, c["H"], drop=O.DropSpec(0.1, 77, 1))
    assert row_max_rel(got, ref) < 2e-2
    assert row_max_rel(got, ref0) > 5e-2            # and the masks really did something
    # loss + gradients under mask set #2
    eng.zero_grad()
    out = eng.forward_loss(tokd, tgtd, memd, None, 0, training=True)
    eng.backward()
    torch.cuda.synchronize()
    spec = O.DropSpec(0.1, *eng.dropout_state())
    assert spec.counter == 2
    lref, g32 = O.loss_and_grads(p, tok, tgt, mem, None, c["H"], drop=spec)
    _, g16 = O.loss_and_grads(p, tok, tgt, mem, None, c["H"], emulate_bf16=True, drop=spec)
    assert abs(out[0].item() - lref.item()) < 1e-3 * lref.item()
    for k in g32:
        got_g = eng.view(k, eng.grads)
        assert _grad_close(got_g, g16[k]), (k, rel_l2(got_g, g16[k]))
        assert _grad_close(got_g, g32[k]), (k, rel_l2(got_g, g32[k]))
    assert float(eng.view("token_embedding.weight", eng.grads)[0].abs().max()) == 0.0
    # fresh masks every training forward; the same (seed, counter) reproduces bit for bit
    l3 = eng.forward_loss(tokd, tgtd, memd, None, 0, training=True)[0].item()
    assert l3 != out[0].item()
    eng.set_dropout(0.1, seed=77)
    eng.forward_logits(tokd, memd, mpm.to(cuda_dev), training=True)
    again = eng.forward_loss(tokd, tgtd, memd, None, 0, training=True)[0].item()
    assert again == out[0].item()


def test_dropout_module_train_eval_and_graph(cuda_dev):
    """nn.Module contract: dropout only in train() mode; a CUDA-graph replay of the fused step draws
    new masks each replay (the counter lives on the device); drop fraction ~ p."""
    from multimodal_image_transformer_b200.decoder import TransformerDecoder
    from multimodal_image_transformer_b200.train import B200AdamW, GraphedTrainStep
    c = CFGS["tiny"]
    torch.manual_seed(5)
    dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.1, pad_idx=0, device=cuda_dev)
    tok, tgt, mem, _ = synth(c, 43)
    tokd, tgtd, memd = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    dec.eval()
    with torch.no_grad():
        a = dec(tokd, memd)
        b = dec(tokd, memd)
    assert torch.equal(a, b)
    dec.train()
    with torch.no_grad():
        l1 = dec.loss(tokd, tgtd, memd)[0].item()
        l2 = dec.loss(tokd, tgtd, memd)[0].item()
    assert l1 != l2
    dec.eval()
    with torch.no_grad():
        assert torch.equal(dec(tokd, memd), a)
    dec.train()
    opt = B200AdamW(dec, lr=1e-4)
    step = GraphedTrainStep(dec, opt, 0, 5.0, warmup=1)
    before = dec.engine.dropout_state()[1]
    losses = [step(memd, tokd, tgtd)[0].item() for _ in range(5)]
    assert dec.engine.dropout_state()[1] == before + 5
    assert len(set(losses)) == 5 and all(torch.isfinite(torch.tensor(losses)))
    # drop fraction of the embedding site, measured through the oracle's restatement of the generator
    spec = O.DropSpec(0.1, 1234, 3)
    frac = 1.0 - spec.keep(0, torch.arange(1 << 20, dtype=torch.int64)).float().mean().item()
    assert abs(frac - 0.1) < 2e-3


def test_projection_path_cfg1(cuda_dev):
    """BASELINE cfg1: 768-wide CLIP features projected to the 512-wide decoder inside the engine."""
    c = CFGS["cfg1"]
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=42)
    tok, tgt, _, _ = synth(c, 43)
    gen = torch.Generator().manual_seed(9)
    feat = torch.randn(c["B"], c["S"], 768, generator=gen)
    lin = torch.nn.Linear(768, c["E"])
    pw, pb = lin.weight.detach().clone(), lin.bias.detach().clone()
    eng = make_engine(c, dict(p, **{"projection.weight": pw, "projection.bias": pb}), cuda_dev, enc_dim=768)
    lref, g16 = O.loss_and_grads(p, tok, tgt, feat, None, c["H"], proj=(pw, pb), emulate_bf16=True)
    with torch.no_grad():
        ref_logits = O.decoder_forward(p, tok, O.project_memory(feat, pw, pb), None, c["H"])
    assert row_max_rel(eng.forward_logits(tok.to(cuda_dev), feat.to(cuda_dev), None), ref_logits) < 2e-2
    eng.zero_grad()
    out = eng.forward_loss(tok.to(cuda_dev), tgt.to(cuda_dev), feat.to(cuda_dev), None, 0, training=True)
    eng.backward()
    assert abs(out[0].item() - lref.item()) < 1e-3 * lref.item()
    for k in ("projection.weight", "projection.bias", "transformer_decoder.layers.0.multihead_attn.in_proj_weight"):
        assert _grad_close(eng.view(k, eng.grads), g16[k]), (k, rel_l2(eng.view(k, eng.grads), g16[k]))


def test_train_trajectory_vs_reference_golden(cuda_dev):
    """3 steps of zero_grad / fwd / CE / bwd / clip(5.0) / AdamW exactly as train.py:80-100."""
    g = load_golden("nano")
    c = g["config"]
    eng = make_engine(c, golden_params(g), cuda_dev)
    tok, tgt, mem = (g[k].to(cuda_dev) for k in ("tokens", "targets", "memory"))
    losses = []
    for _ in range(3):
        eng.zero_grad()
        out = eng.forward_loss(tok, tgt, mem, None, 0, training=True)
        eng.backward()
        eng.adamw_step(lr=g["train_lr"], betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5, max_norm=5.0)
        losses.append(out[0].item())
    for a, b in zip(losses, g["train_losses"]):
        assert abs(a - b) < 2e-3 * b, (losses, g["train_losses"])
    for k, v in g["params_after"].items():
        d = (eng.view(k).cpu() - v).abs()
        assert d.max() <= 2 * 3 * g["train_lr"] + 1e-6, k        # Adam sign sensitivity, see test_oracle.py
        if "in_proj_bias" not in k:                              # key-bias gradients are exactly zero: pure sign noise
            assert d.mean() < 0.35 * g["train_lr"], k


def test_graphed_train_step_matches_reference_golden(cuda_dev):
    """train.GraphedTrainStep (eager warm-up step, capture, replay) follows the same trajectory."""
    from multimodal_image_transformer_b200.decoder import TransformerDecoder
    from multimodal_image_transformer_b200.train import B200AdamW, GraphedTrainStep
    g = load_golden("nano")
    c = g["config"]
    torch.manual_seed(g["seed"])
    dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0).train()
    opt = B200AdamW(dec, lr=g["train_lr"], betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    step = GraphedTrainStep(dec, opt, 0, 5.0, warmup=1)
    tok, tgt, mem = (g[k].to(cuda_dev) for k in ("tokens", "targets", "memory"))
    losses = [step(mem, tok, tgt)[0].item() for _ in range(3)]
    assert step.graph is not None
    for a, b in zip(losses, g["train_losses"]):
        assert abs(a - b) < 2e-3 * b, (losses, g["train_losses"])
    assert dec.engine.opt_step == 3 and int(dec.engine._step_dev.item()) == 3


def test_scheduler_drives_the_device_learning_rate_also_under_graph_replay(cuda_dev):
    """A LambdaLR over B200AdamW (reference train.py:331-341) moves param_groups[0]["lr"]; the fused kernel reads the
    learning rate from device memory, eagerly and when the step is replayed from a CUDA graph."""
    from multimodal_image_transformer_b200.decoder import TransformerDecoder
    from multimodal_image_transformer_b200.train import B200AdamW, GraphedTrainStep
    c = CFGS["nano"]
    torch.manual_seed(1)
    dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0, device=cuda_dev).train()
    opt = B200AdamW(dec, lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: min(1.0, (s + 1) / 4.0))
    step = GraphedTrainStep(dec, opt, 0, 5.0, warmup=1)
    tok, tgt, mem, _ = synth(c, 43)
    tokd, tgtd, memd = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    seen, before = [], dec.engine.params.clone()
    for i in range(6):
        want = sched.get_last_lr()[0]
        step(memd, tokd, tgtd)
        seen.append((want, float(dec.engine._lr_dev.item())))
        sched.step()
    assert step.graph is not None
    for want, got in seen:
        assert abs(want - got) <= 1e-9 + 1e-6 * want, seen
    assert seen[0][0] == pytest.approx(2.5e-4) and seen[-1][0] == pytest.approx(1e-3)
    assert not torch.equal(before, dec.engine.params)


def test_captured_graph_pins_the_workspace(cuda_dev):
    """A captured train step holds raw pointers into the activation workspace: a later forward that would force a
    reallocation must fail loudly instead of letting the next replay run on freed memory; reserving the larger
    shape before capture makes the same sequence legal."""
    from multimodal_image_transformer_b200.decoder import TransformerDecoder
    from multimodal_image_transformer_b200.train import B200AdamW, GraphedTrainStep
    c = CFGS["nano"]
    tok, tgt, mem, _ = synth(c, 43)
    tokd, tgtd, memd = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    big = (tokd.repeat(8, 1), memd.repeat(8, 1, 1))
    for reserve in (False, True):
        torch.manual_seed(1)
        dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0, device=cuda_dev).train()
        opt = B200AdamW(dec, lr=1e-3)
        if reserve:
            dec.engine.reserve_workspace(big[0].shape[0], big[0].shape[1], big[1].shape[1], big[1].shape[2], False)
        step = GraphedTrainStep(dec, opt, 0, 5.0, warmup=1)
        for _ in range(3):
            step(memd, tokd, tgtd)
        assert step.graph is not None
        if reserve:
            with torch.no_grad():
                dec.eval()(big[0], big[1])
            dec.train()
            assert torch.isfinite(step(memd, tokd, tgtd)).all()
        else:
            with pytest.raises(RuntimeError, match="reserve_workspace"):
                with torch.no_grad():
                    dec.eval()(big[0], big[1])
            step.release()
            with torch.no_grad():
                dec.eval()(big[0], big[1])             # legal again once the graph is gone


def test_shadow_weights_follow_p_data_edits_outside_the_fused_loop(cuda_dev):
    """Edits through `p.data` bump no version counter: eval-mode forwards, generation and the first
    forward after a train()/eval() switch re-cast the bf16 shadow, so they never run on stale weights."""
    from multimodal_image_transformer_b200.decoder import TransformerDecoder
    c = CFGS["nano"]
    torch.manual_seed(1)
    dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0, device=cuda_dev).eval()
    tok, tgt, mem, _ = synth(c, 43)
    tokd, memd = tok.to(cuda_dev), mem.to(cuda_dev)
    with torch.no_grad():
        a = dec(tokd, memd).clone()
        orig = dec.fc_out.bias.data.clone()
        dec.fc_out.bias.data.add_(1.0)                                   # invisible to _version
        b = dec(tokd, memd)
    assert (b - a - 1.0).abs().max().item() < 2e-2
    with torch.no_grad():
        dec.fc_out.bias.data.copy_(orig)
        dec.engine.decode_begin(memd, None, beam=1, max_len=6)            # generation re-casts too
        assert torch.equal(dec(tokd, memd), a)


def test_dropin_module_state_dict_and_autograd(cuda_dev):
    """decoder.TransformerDecoder: reference constructor signature, bit-identical seeded init,
    reference state_dict keys/shapes, and the logits -> criterion -> backward() loop of train.py."""
    from multimodal_image_transformer_b200.decoder import TransformerDecoder
    c = CFGS["nano"]
    torch.manual_seed(42)
    dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=42)
    sd = dec.state_dict()
    assert list(sd.keys()) == list(p.keys()) or set(sd.keys()) == set(p.keys())
    for k in p:
        assert sd[k].shape == p[k].shape and torch.equal(sd[k].cpu(), p[k]), k
    tok, tgt, mem, mpm = synth(c, 43)
    dec.eval()
    with torch.no_grad():
        lg = dec(tok.to(cuda_dev), mem.to(cuda_dev), mpm.to(cuda_dev))
        ref = O.decoder_forward(p, tok, mem, mpm, c["H"])
    assert lg.shape == (c["B"], c["T"], c["V"]) and row_max_rel(lg, ref) < 2e-2
    dec.train()
    logits = dec(tok.to(cuda_dev), mem.to(cuda_dev))
    loss = torch.nn.CrossEntropyLoss(ignore_index=0)(logits.view(-1, c["V"]), tgt.to(cuda_dev).reshape(-1))
    for q in dec.parameters():
        q.grad = None
    loss.backward()
    _, g16 = O.loss_and_grads(p, tok, tgt, mem, None, c["H"], emulate_bf16=True)
    named = dict(dec.named_parameters())
    assert set(named) == set(g16)
    for k, v in g16.items():
        assert _grad_close(named[k].grad, v), (k, rel_l2(named[k].grad, v))
    # load_state_dict round trip with perturbed reference-format tensors
    sd2 = {k: (v + 0.01 if v.is_floating_point() and k != "positional_encoding.pe" else v) for k, v in p.items()}
    dec.load_state_dict(sd2, strict=True)
    with torch.no_grad():
        lg2 = dec.eval()(tok.to(cuda_dev), mem.to(cuda_dev))
        ref2 = O.decoder_forward(sd2, tok, mem, None, c["H"])
    assert row_max_rel(lg2, ref2) < 2e-2


def test_greedy_vs_reference_golden_and_oracle(cuda_dev):
    from multimodal_image_transformer_b200.decoder import TransformerDecoder  # noqa: F401
    # (the cfg2 golden is not used here: with random-init weights its four 12-token captions contain a bf16 near-tie
    # of the arg-max; the forward / loss test above pins that shape)
    for name in ("nano", "cfg1"):
        g = load_golden(name)
        c = g["config"]
        eng = make_engine(c, golden_params(g), cuda_dev)
        n = len(g["greedy"])
        eng.decode_begin(g["memory"][:n].to(cuda_dev), None, beam=1, max_len=g["greedy_max_len"])
        toks, lens = eng.generate_greedy(1, 2, g["greedy_max_len"], stop_check_interval=0)
        got = [toks[b, :int(lens[b])].tolist() for b in range(n)]
        same = sum(a == b for a, b in zip(got, g["greedy"]))
        assert same >= 0.99 * n, (got, g["greedy"])


def _teacher_forced_check(p, c, mem, toks, lens, end_id=2, tie=2e-2):
    """Greedy decisions of the GPU against the fp32 oracle, teacher-forced: ONE oracle forward over the GPU's own
    sequences gives the reference logits of every step (causal + key-padding masks make position t depend on the
    prefix only: SURVEY appendix A, KV cache == full recompute).  Returns (decisions, agreeing, unexplained):
    a disagreement is `explained` iff the oracle's top-2 margin at that step is below tie * max|logit| of the row
    (a bf16 near-tie of the arg-max)."""
    toks, lens = toks.cpu(), lens.cpu()
    B, T = toks.shape
    with torch.no_grad():
        ref = O.decoder_forward(p, toks[:, :T - 1], mem, None, c["H"])            # logits of steps 0 .. T-2
    n = agree = unexplained = 0
    for b in range(B):
        for t in range(int(lens[b]) - 1):                     # step t emitted toks[b, t + 1]
            row = ref[b, t]
            got, want = int(toks[b, t + 1]), int(row.argmax())
            n += 1
            if got == want:
                agree += 1
                continue
            top2 = row.topk(2).values
            if not (float(top2[0] - row[got]) < tie * float(row.abs().max())):
                unexplained += 1
    return n, agree, unexplained


def test_greedy_cfg2_reference_golden_with_near_tie_rule(cuda_dev):
    """The greedy fixture of the unmodified reference at the benchmark decoder shape (cfg2, 4 images x 12 tokens).
    Random-init logits are nearly flat, so a bf16 near-tie can flip an arg-max and everything after it: sequences
    must be identical up to the first flip, and that flip must be a near-tie by the fp32 oracle's own margin."""
    g = load_golden("cfg2")
    c = g["config"]
    p = golden_params(g)
    eng = make_engine(c, p, cuda_dev)
    n = len(g["greedy"])
    mem = g["memory"][:n]
    eng.decode_begin(mem.to(cuda_dev), None, beam=1, max_len=g["greedy_max_len"])
    toks, lens = eng.generate_greedy(1, 2, g["greedy_max_len"], stop_check_interval=0)
    got = [toks[b, :int(lens[b])].tolist() for b in range(n)]
    for b, (a, r) in enumerate(zip(got, g["greedy"])):
        if a == r:
            continue
        t = next(i for i in range(min(len(a), len(r))) if a[i] != r[i])         # first differing token (emitted at step t-1)
        with torch.no_grad():
            row = O.decoder_forward(p, torch.tensor([r[:t]]), mem[b:b + 1], None, c["H"])[0, -1]
        assert int(row.argmax()) == r[t]                                          # the oracle reproduces the reference
        assert float(row.max() - row[a[t]]) < 2e-2 * float(row.abs().max()), (b, t, a, r)
    steps, agree, unexplained = _teacher_forced_check(p, c, mem, toks, lens)
    assert unexplained == 0 and agree >= 0.95 * steps, (steps, agree, unexplained)


def test_greedy_cfg4_shape_teacher_forced(cuda_dev):
    """BASELINE configs[3] decoder shape (E=768, H=12, L=6, F=3072, S=197; the shape the captions/s number is quoted on)
    over 128 images x 12 tokens: every greedy decision equals the fp32 oracle's arg-max given the same prefix, or is a
    near-tie by the oracle's margin (2e-2 of the row's largest |logit|).  Random-init logits are nearly flat, so such
    ties are frequent: measured on the B200, 97.9 % of the 1408 decisions are identical and the remaining 2.1 % are all
    ties; asserted: every decision is identical or a tie (north_star's ">= 99 % identical" reads on this sum), and
    >= 96 % are identical outright."""
    c = dict(V=10000, E=768, H=12, L=6, F=3072, ML=48, B=128, T=12, S=197)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=42)
    mem = torch.randn(c["B"], c["S"], c["E"], generator=torch.Generator().manual_seed(44))
    eng = make_engine(c, p, cuda_dev)
    eng.decode_begin(mem.to(cuda_dev), None, beam=1, max_len=c["T"])
    toks, lens = eng.generate_greedy(1, 2, c["T"], stop_check_interval=0)
    steps, agree, unexplained = _teacher_forced_check(p, c, mem, toks, lens)
    assert steps >= c["B"] * (c["T"] - 1) * 0.5
    assert unexplained == 0, (steps, agree, unexplained)
    assert agree >= 0.96 * steps, (steps, agree, unexplained)


def _oracle_seq_score(p, c, mem_b, seq, end_id=2):
    """Sum of token log-probabilities of `seq` (START first) under the fp32 oracle, up to and including its first END."""
    with torch.no_grad():
        lp = torch.log_softmax(O.decoder_forward(p, torch.tensor([seq[:-1]]), mem_b, None, c["H"])[0], dim=-1)
    return float(sum(lp[t, seq[t + 1]] for t in range(len(seq) - 1)))


@pytest.mark.parametrize("beam", [1, 3])
def test_generated_pad_is_masked_as_key(cuda_dev, beam):
    """A PAD id emitted before END is masked as a self-attention key from then on: the reference rebuilds
    tgt_key_padding_mask = (tokens == pad_idx) from the whole prefix on every generate() step (decoder.py:162,
    model.py:224-228).  fc_out.bias makes PAD the arg-max at many steps (14 of 16 oracle captions contain a PAD
    followed by real tokens); the KV-cached path must take the oracle's decisions given the same prefix."""
    c = dict(CFGS["tiny"], B=16)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=3)
    p["fc_out.bias"][0] += 1.3
    p["fc_out.bias"][2] += 0.3
    mem = torch.randn(c["B"], c["S"], c["E"], generator=torch.Generator().manual_seed(4))
    eng = make_engine(c, p, cuda_dev)
    eng.decode_begin(mem.to(cuda_dev), None, beam=beam, max_len=12)
    if beam == 1:
        toks, lens = eng.generate_greedy(1, 2, 12, stop_check_interval=0)
        got = [toks[b, :int(lens[b])].tolist() for b in range(c["B"])]
        assert sum(any(t != 0 for t in r[r.index(0):]) for r in got if 0 in r[1:]) >= 6, got      # the scenario occurs
        steps, agree, unexplained = _teacher_forced_check(p, c, mem, toks, lens)
        assert unexplained == 0 and agree >= 0.9 * steps, (steps, agree, unexplained)
        # an oracle WITHOUT the key-padding mask disagrees: the check above is sensitive to the mask
        with torch.no_grad():
            ref_nomask = O.decoder_forward(p, toks.cpu()[:, :-1], mem, None, c["H"], pad_idx=-1)
        t_ = toks.cpu()
        flips = sum(int(ref_nomask[b, t].argmax()) != int(t_[b, t + 1]) for b in range(c["B"]) for t in range(int(lens[b]) - 1))
        assert flips > steps - agree
        # step API: the engine remembers the prefix (decode_step sees one column at a time)
        eng.decode_begin(mem.to(cuda_dev), None, beam=1, max_len=12)
        cur = torch.full((c["B"],), 1, dtype=torch.int64, device=cuda_dev)
        cols = [cur.cpu()]
        for t in range(11):
            cur = eng.decode_step(cur, t)
            cols.append(cur.cpu())
        stepped = torch.stack(cols, 1)
        for b in range(c["B"]):
            n_ = int(lens[b])
            assert stepped[b, :n_].tolist() == got[b], (b, stepped[b].tolist(), got[b])
    else:
        toks, lens, _ = eng.generate_beam(1, 2, 12)
        got = [toks[b, :int(lens[b])].tolist() for b in range(c["B"])]
        with torch.no_grad():
            ref = O.beam_generate(p, mem, 1, 2, 12, c["H"], beam_size=beam)
        assert sum(0 in r[1:] for r in ref) >= 6, ref
        # PAD-heavy, nearly flat distributions: a bf16 near-tie at one pruning step can drop the hypothesis the fp32 search
        # ends up preferring, so sequences are compared by their ORACLE score: never much worse than the oracle's own
        # best (0.5 nat on a ~-57 nat sequence), close on average, and mostly identical
        same, gaps = 0, []
        for b in range(c["B"]):
            if got[b] == ref[b]:
                same += 1
                gaps.append(0.0)
                continue
            gaps.append(_oracle_seq_score(p, c, mem[b:b + 1], ref[b]) - _oracle_seq_score(p, c, mem[b:b + 1], got[b]))
        assert max(gaps) < 0.5 and sum(gaps) / len(gaps) < 0.1, (same, gaps)
        assert same >= c["B"] // 2, (same, gaps)


def test_greedy_batch_vs_oracle_and_early_stop(cuda_dev):
    c = dict(CFGS["tiny"], B=24)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=3)
    p["fc_out.bias"][2] += 1.5          # make END likely enough that some rows stop early
    mem = torch.randn(c["B"], c["S"], c["E"], generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        ref = O.greedy_generate(p, mem, 1, 2, 14, c["H"])
    eng = make_engine(c, p, cuda_dev)
    eng.decode_begin(mem.to(cuda_dev), None, beam=1, max_len=14)
    for interval in (0, 4):
        toks, lens = eng.generate_greedy(1, 2, 14, stop_check_interval=interval)
        got = [toks[b, :int(lens[b])].tolist() for b in range(c["B"])]
        same = sum(a == b for a, b in zip(got, ref))
        assert same >= 0.95 * c["B"], (same, got[:3], ref[:3])     # 24 rows: allow one bf16 near-tie
        for b in range(c["B"]):
            assert toks[b, int(lens[b]):].eq(0).all()              # PAD after END
        eng.decode_begin(mem.to(cuda_dev), None, beam=1, max_len=14)


def test_kv_cache_step_equals_full_recompute(cuda_dev):
    """decode_step logits-argmax at position t == argmax of the full forward on the prefix."""
    c, p, (tok, _, mem, mpm) = _case("tiny")
    tok = tok.clone()
    tok[tok == 0] = 7
    eng = make_engine(c, p, cuda_dev)
    full = eng.forward_logits(tok.to(cuda_dev), mem.to(cuda_dev), mpm.to(cuda_dev)).argmax(-1).cpu()
    eng.decode_begin(mem.to(cuda_dev), mpm.to(cuda_dev), beam=1, max_len=c["T"] + 1)
    agree = 0
    for t in range(c["T"]):
        nxt = eng.decode_step(tok[:, t].to(cuda_dev), t).cpu()
        agree += int((nxt == full[:, t]).sum())
    assert agree >= 0.98 * tok.numel()


def test_beam_search_vs_oracle_spec(cuda_dev):
    c = dict(CFGS["nano"], B=6)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=5)
    p["fc_out.bias"][2] += 1.0
    mem = torch.randn(c["B"], c["S"], c["E"], generator=torch.Generator().manual_seed(6))
    eng = make_engine(c, p, cuda_dev)
    with torch.no_grad():
        ref3 = O.beam_generate(p, mem, 1, 2, 10, c["H"], beam_size=3)
        greedy = O.greedy_generate(p, mem, 1, 2, 10, c["H"])
    eng.decode_begin(mem.to(cuda_dev), None, beam=3, max_len=10)
    toks, lens, score = eng.generate_beam(1, 2, 10)
    got = [toks[b, :int(lens[b])].tolist() for b in range(c["B"])]
    assert sum(a == b for a, b in zip(got, ref3)) >= c["B"] - 1, (got, ref3)
    assert torch.isfinite(score).all()
    # beam = 1 through the model-level API degenerates to greedy
    from multimodal_image_transformer_b200.model import generate_from_memory

    class _D:      # minimal stand-in exposing .engine
        engine = eng
    out = generate_from_memory(_D, mem.to(cuda_dev), 1, 2, 10, method="beam", beam_size=1)
    assert sum(a == b for a, b in zip(out, greedy)) >= c["B"] - 1


@pytest.mark.parametrize("name,beam", [("tiny", 1), ("hd96", 1), ("tiny", 3)])
def test_decode_partitions_are_equivalent(cuda_dev, name, beam, monkeypatch):
    """Generation over independent image partitions on parallel streams (graph replay included) gives
    exactly the tokens of the single-partition run: every decode op is row-local."""
    c = dict(CFGS[name], B=11)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=8)
    p["fc_out.bias"][2] += 1.0
    g = torch.Generator().manual_seed(9)
    mem = torch.randn(c["B"], c["S"], c["E"], generator=g)
    mpm = torch.zeros(c["B"], c["S"], dtype=torch.bool)
    mpm[2, c["S"] // 3:] = True
    eng = make_engine(c, p, cuda_dev)
    outs = []
    for parts in ("1", "3", "4"):
        monkeypatch.setenv("B200_DECODE_PARTS", parts)
        runs = []
        for rep in range(3):      # eager, capture, replay
            eng.decode_begin(mem.to(cuda_dev), mpm.to(cuda_dev), beam=beam, max_len=12)
            if beam == 1:
                toks, lens = eng.generate_greedy(1, 2, 12, stop_check_interval=0)
            else:
                toks, lens, _ = eng.generate_beam(1, 2, 12)
            runs.append((toks.cpu().clone(), lens.cpu().clone()))
        for t, l in runs[1:]:
            assert torch.equal(t, runs[0][0]) and torch.equal(l, runs[0][1])
        outs.append(runs[0])
    for t, l in outs[1:]:
        assert torch.equal(t, outs[0][0]) and torch.equal(l, outs[0][1])

def test_trim_batch_keeps_loss_and_gradients(cuda_dev):
    """train.trim_batch drops the all-PAD tail columns of a collated batch (the reference pads every
    caption to MAX_SEQ_LEN): same loss, same gradients, less work."""
    from multimodal_image_transformer_b200.train import trim_batch
    c = dict(CFGS["tiny"], T=40)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=11)
    tok, tgt, mem, _ = synth(c, 12)
    tok[:, 19:] = 0
    tgt[:, 18:] = 0                         # longest caption: 19 tokens -> 24 kept columns
    tok2, tgt2 = trim_batch(tok, tgt, 0)
    assert tok2.shape[1] == 24 and torch.equal(tok2, tok[:, :24])
    eng = make_engine(c, p, cuda_dev)
    res = []
    for a, b in ((tok, tgt), (tok2, tgt2)):
        eng.zero_grad()
        out = eng.forward_loss(a.to(cuda_dev), b.to(cuda_dev), mem.to(cuda_dev), None, 0, training=True)
        eng.backward()
        res.append((out.cpu().clone(), eng.grads.clone()))
    assert res[0][0][1].item() == res[1][0][1].item()
    assert abs(res[0][0][0].item() - res[1][0][0].item()) < 1e-6 * res[0][0][0].item()
    assert rel_l2(res[1][1], res[0][1]) < 2e-3          # fp32 atomics / split-K order only


def test_optimizer_state_dict_is_torch_adamw_shaped(cuda_dev):
    """B200AdamW.state_dict() has the layout of torch.optim.AdamW over model.parameters() (reference
    train.py:319-325, saved at :424, restored at :351-357): one fused step here equals one torch
    AdamW step on the same gradients, entry by entry, and the dict round-trips."""
    from multimodal_image_transformer_b200.decoder import TransformerDecoder
    from multimodal_image_transformer_b200.train import B200AdamW
    c = CFGS["nano"]
    torch.manual_seed(3)
    dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0, device=cuda_dev)
    dec.train()
    opt = B200AdamW(dec, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    names = list(dec.engine.layout)
    before = {n: dec.engine.view(n).detach().cpu().clone() for n in names}
    tok, tgt, mem, _ = synth(c, 5)
    opt.zero_grad()
    dec.loss(tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev))
    dec.backward()
    grads = {n: dec.engine.view(n, dec.engine.grads).detach().cpu().clone() for n in names}
    opt.step(max_grad_norm=0.0)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(len(names)))
    ref_params = [torch.nn.Parameter(before[n].clone()) for n in names]
    ref = torch.optim.AdamW(ref_params, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    for p_, n in zip(ref_params, names):
        p_.grad = grads[n].clone()
    ref.step()
    rsd = ref.state_dict()
    for i, n in enumerate(names):
        assert float(sd["state"][i]["step"]) == float(rsd["state"][i]["step"]) == 1.0
        assert rel_l2(sd["state"][i]["exp_avg"], rsd["state"][i]["exp_avg"]) < 1e-5, n
        assert rel_l2(sd["state"][i]["exp_avg_sq"], rsd["state"][i]["exp_avg_sq"]) < 1e-5, n
        assert (dec.engine.view(n).cpu() - ref_params[i].detach()).abs().max() < 1e-6, n
    # a reference-written optimizer checkpoint restores into the fused optimizer (and ours into torch)
    opt2 = B200AdamW(dec, lr=5e-4)
    opt2.load_state_dict(rsd)
    assert dec.engine.opt_step == 1 and opt2.param_groups[0]["lr"] == 1e-3
    assert rel_l2(dec.engine.exp_avg, opt.engine.exp_avg) < 1e-6
    ref.load_state_dict(sd)


def test_full_size_properties_cfg2(cuda_dev):
    """BASELINE cfg2 (B=256, T=47, S=197, E=768, H=12, F=3072, L=6, V=10000): too big for the CPU
    oracle, so check size-independent properties: loss at init ~ ln V, gradient norms finite, padding
    row untouched, batch-permutation invariance of the loss, and equality of the fused loss with
    the loss computed from materialised logits."""
    c = dict(V=10000, E=768, H=12, L=6, F=3072, ML=100, B=256, T=47, S=197)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=42)
    tok, tgt, mem, _ = synth(c, 43)
    eng = make_engine(c, p, cuda_dev)
    tokd, tgtd, memd = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    eng.zero_grad()
    out = eng.forward_loss(tokd, tgtd, memd, None, 0, training=True)
    eng.backward()
    loss = out[0].item()
    assert abs(loss - 9.21) < 0.35 and out[1].item() == (tgt != 0).sum().item()
    assert torch.isfinite(eng.grads).all() and eng.grads.norm().item() > 0
    assert float(eng.view("token_embedding.weight", eng.grads)[0].abs().max()) == 0.0
    logits = eng.forward_logits(tokd, memd, None)
    lref = torch.nn.functional.cross_entropy(logits.view(-1, c["V"]), tgtd.view(-1), ignore_index=0).item()
    assert abs(loss - lref) < 2e-4 * lref
    perm = torch.randperm(c["B"], generator=torch.Generator().manual_seed(1))
    out_p = eng.forward_loss(tokd[perm], tgtd[perm], memd[perm], None, 0, training=False)
    assert abs(out_p[0].item() - loss) < 1e-5 * loss


# ------------------------------------------------------------------ generation schedules / side streams
_KNOBS = ("B200_DECODE_PARTS", "B200_DEC_ATTN_GRID", "B200_DEC_KV_FLAGS", "B200_DEC_ATTN_DYN", "B200_DEC_ATTN_STREAM",
          "B200_DEC_PREFETCH_MB", "B200_DEC_SINGLE_CTA", "B200_DEC_KSPLIT_E", "B200_DEC_KSPLIT_F", "B200_DEC_GEMM_CTAS")


@pytest.mark.parametrize("heads,beam", [(4, 1), (2, 1), (4, 3)])     # hd = 64 / 128: the tensor-core decode attention
def test_decode_attention_schedules_are_equivalent(cuda_dev, heads, beam, monkeypatch):
    """The cross-attention launch shapes of generation -- small CTAs on every SM, fat multi-unit CTAs on an SM
    budget, 3-deep rings, 16-row tail boxes, eviction hints, L2 warm-up, dynamic item distribution, the dedicated
    attention stream -- only move the same arithmetic around: token ids are identical to the plain schedule's."""
    c = dict(V=1000, E=256, H=heads, L=2, F=512, ML=40, B=37, T=17, S=197)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=5)
    g = torch.Generator().manual_seed(6)
    mem = torch.randn(c["B"], c["S"], c["E"], generator=g)
    mpm = torch.zeros(c["B"], c["S"], dtype=torch.bool)
    mpm[3, 150:] = True
    eng = make_engine(c, p, cuda_dev)

    def run(cfg):
        for k in _KNOBS:
            monkeypatch.delenv(k, raising=False)
        for k, v in cfg.items():
            monkeypatch.setenv(k, str(v))
        outs = []
        for rep in range(3):      # eager, capture, replay
            eng.decode_begin(mem.to(cuda_dev), mpm.to(cuda_dev), beam=beam, max_len=10)
            if beam == 1:
                toks, lens = eng.generate_greedy(1, 2, 10, stop_check_interval=0)
            else:
                toks, lens, _ = eng.generate_beam(1, 2, 10)
            outs.append((toks.cpu().clone(), lens.cpu().clone()))
        for t, l in outs[1:]:
            assert torch.equal(t, outs[0][0]) and torch.equal(l, outs[0][1]), cfg
        return outs[0]

    plain = {"B200_DECODE_PARTS": 1, "B200_DEC_KV_FLAGS": 0, "B200_DEC_SINGLE_CTA": 0}
    base = run(plain)
    variants = [
        dict(plain, B200_DEC_KV_FLAGS=3),
        dict(plain, B200_DEC_ATTN_GRID=7),                                        # fat CTAs, one partition
        dict(plain, B200_DEC_ATTN_GRID=7, B200_DEC_KV_FLAGS=11),                 # + 3-deep rings (greedy, hd 64)
        dict(plain, B200_DEC_ATTN_DYN=1),
        dict(plain, B200_DEC_ATTN_GRID=5, B200_DEC_ATTN_DYN=1, B200_DEC_KV_FLAGS=2),
        dict(plain, B200_DEC_PREFETCH_MB=1),
        dict(plain, B200_DECODE_PARTS=3, B200_DEC_ATTN_GRID=9, B200_DEC_ATTN_STREAM=1),
        dict(plain, B200_DECODE_PARTS=4, B200_DEC_ATTN_GRID=6, B200_DEC_ATTN_DYN=1, B200_DEC_ATTN_STREAM=1, B200_DEC_KV_FLAGS=15),
        {},                                                                       # the shipped defaults
    ]
    for cfg in variants:
        t, l = run(cfg)
        assert torch.equal(t, base[0]) and torch.equal(l, base[1]), cfg


def test_decode_plan_info_reports_the_schedule(cuda_dev, monkeypatch):
    """b200_engine_decode_plan_info: greedy generation over a stream-heavy batch runs as four concurrent partitions
    with the cross attention on 70 % of the SMs; beam search and small batches keep one partition."""
    for k in _KNOBS:
        monkeypatch.delenv(k, raising=False)
    c = dict(V=264, E=768, H=12, L=1, F=256, ML=16, B=256, T=8, S=197)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=3)
    eng = make_engine(c, p, cuda_dev)
    mem = torch.randn(c["B"], c["S"], c["E"], device=cuda_dev)
    eng.decode_begin(mem, None, beam=1, max_len=8)
    info = eng.decode_plan_info()
    sms = torch.cuda.get_device_properties(cuda_dev).multi_processor_count
    assert info["partitions"] == 4 and info["attention_sms"] == (sms * 70 + 50) // 100
    assert info["gemm_grid_cap"] == sms - info["attention_sms"] and info["split_k_e"] == 2 and info["split_k_f"] == 1
    toks, lens = eng.generate_greedy(1, 2, 8, 0)
    assert toks.shape == (256, 8) and int(lens.min()) >= 1
    eng.decode_begin(mem, None, beam=2, max_len=8)
    assert eng.decode_plan_info()["partitions"] == 1
    eng.decode_begin(mem[:8], None, beam=1, max_len=8)
    info = eng.decode_plan_info()
    assert info["partitions"] == 1 and info["attention_sms"] == 0 and info["gemm_grid_cap"] == 0


@pytest.mark.parametrize("name", ["tiny", "cfg1"])
def test_backward_bias_side_stream_matches_inline(cuda_dev, name, monkeypatch):
    """The bias-gradient column sums run on a side stream next to the backward GEMMs (which give up a pipeline
    stage); the gradients are those of the single-stream backward (atomics: equal up to summation order)."""
    c, p, (tok, tgt, mem, _) = _case(name)
    eng = make_engine(c, p, cuda_dev)
    grads = []
    for inline in (True, False, False):
        if inline:
            monkeypatch.setenv("B200_BIAS_INLINE", "1")
        else:
            monkeypatch.delenv("B200_BIAS_INLINE", raising=False)
        eng.zero_grad()
        eng.forward_loss(tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev), None, 0, training=True)
        eng.backward()
        torch.cuda.synchronize()
        grads.append(eng.grads.clone())
    for g in grads[1:]:
        assert rel_l2(g, grads[0]) < 1e-5
        for k in ("fc_out.bias", "transformer_decoder.layers.0.linear1.bias", "transformer_decoder.layers.0.self_attn.in_proj_bias",
                  "transformer_decoder.layers.1.multihead_attn.in_proj_bias"):
            assert rel_l2(eng.view(k, g), eng.view(k, grads[0])) < 1e-5, k


def _varlen_case(name, seed=7, lo=12):
    """A batch whose captions have U[lo, T] real tokens followed by PAD (the reference's padded batches,
    tokenizer.py:293-313 / dataset.py:176-206)."""
    c = CFGS[name]
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    B, T, S, V = c["B"], c["T"], c["S"], c["V"]
    cap = torch.randint(4, V, (B, T + 1), generator=g)
    cap[:, 0] = 1
    lens = torch.randint(min(lo, T), T + 1, (B,), generator=g)          # tokens of the INPUT (caption[:-1]) that are real
    lens[0] = T                                                         # one full-length caption
    for b in range(B):
        n = int(lens[b])
        if n < T + 1:
            cap[b, n] = 2 if n < T + 1 else cap[b, n]                  # END closes the caption ...
            cap[b, n + 1:] = 0                                          # ... PAD after it
    tok, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
    mem = torch.randn(B, S, c["E"], generator=g)
    return c, p, tok, tgt, mem


@pytest.mark.parametrize("name", ["nano", "hd96", "cfg1", "cfg2s"])
def test_packed_varlen_path_equals_padded_path(cuda_dev, name):
    """f4: the packed / var-len engine path (cu_seqlens, rows = sum of lengths) gives the padded path's
    loss and gradients on captions of U[12, T] tokens, and both match the fp32 oracle (= the reference's padded math)."""
    from multimodal_image_transformer_b200.engine import DecoderEngine
    c, p, tok, tgt, mem = _varlen_case(name)
    lengths = DecoderEngine.packed_lengths(tok, 0)
    assert lengths is not None and int(lengths.sum()) < tok.numel()
    eng = make_engine(c, p, cuda_dev)
    td, gd, md = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    eng.zero_grad()
    out_pad = eng.forward_loss(td, gd, md, None, 0, training=True).cpu()
    eng.backward()
    g_pad = eng.grads.clone()
    eng.zero_grad()
    out_pk = eng.forward_loss(td, gd, md, None, 0, training=True, lengths=lengths).cpu()
    eng.backward()
    g_pk = eng.grads.clone()
    torch.cuda.synchronize()
    assert out_pk[1].item() == out_pad[1].item() == (tgt != 0).sum().item()
    # same kernels on the same rows (only the attention tiling differs): loss to 1e-6 relative, gradients to 2e-3
    assert abs(out_pk[0].item() - out_pad[0].item()) <= 1e-6 * abs(out_pad[0].item()) + 1e-7
    assert rel_l2(g_pk, g_pad) < 2e-3
    for k in ("fc_out.weight", "token_embedding.weight", "transformer_decoder.layers.0.self_attn.in_proj_weight",
              "transformer_decoder.layers.0.multihead_attn.in_proj_weight", "transformer_decoder.layers.0.linear1.weight"):
        assert rel_l2(eng.view(k, g_pk), eng.view(k, g_pad)) < 2e-3, k
    # and against the oracle (the reference's padded computation)
    lref, gref = O.loss_and_grads(p, tok, tgt, mem, None, c["H"])
    assert abs(out_pk[0].item() - lref.item()) < 1e-3 * lref.item()
    assert _grad_close(eng.view("fc_out.weight", g_pk), gref["fc_out.weight"], rel=1.5e-2, cos=0.9995)
    assert _grad_close(eng.view("token_embedding.weight", g_pk), gref["token_embedding.weight"])


def test_packed_lengths_rejects_inner_pad():
    from multimodal_image_transformer_b200.engine import DecoderEngine
    tok = torch.tensor([[1, 5, 6, 2, 0, 0], [1, 7, 0, 8, 2, 0]])
    assert DecoderEngine.packed_lengths(tok, 0) is None
    assert DecoderEngine.packed_lengths(tok[:1], 0).tolist() == [4]


def test_packed_path_with_dropout_and_eval(cuda_dev):
    """The packed path in training mode with dropout (finite loss / gradients; masks are drawn per packed row, so only
    the dropout-free quantities are compared) and in eval mode (identical to the padded eval loss)."""
    from multimodal_image_transformer_b200.engine import DecoderEngine
    c, p, tok, tgt, mem = _varlen_case("tiny", seed=11, lo=5)
    lengths = DecoderEngine.packed_lengths(tok, 0)
    eng = make_engine(c, p, cuda_dev)
    td, gd, md = tok.to(cuda_dev), tgt.to(cuda_dev), mem.to(cuda_dev)
    ev_pad = eng.forward_loss(td, gd, md, None, 0, training=False).cpu()
    ev_pk = eng.forward_loss(td, gd, md, None, 0, training=False, lengths=lengths).cpu()
    assert abs(ev_pk[0].item() - ev_pad[0].item()) <= 1e-6 * abs(ev_pad[0].item()) + 1e-7
    eng.set_dropout(0.1, seed=3)
    eng.dropout_active(True)
    eng.zero_grad()
    out = eng.forward_loss(td, gd, md, None, 0, training=True, lengths=lengths).cpu()
    eng.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and torch.isfinite(eng.grads).all()
    assert abs(out[0].item() - ev_pad[0].item()) < 0.2 * ev_pad[0].item()      # dropout moves the loss, not its scale
    assert out[1].item() == (tgt != 0).sum().item()
    eng.dropout_active(False)
