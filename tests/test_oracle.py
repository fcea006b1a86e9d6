"""Pins the CPU oracle: against the committed golden vectors (produced by the unmodified reference,
tests/golden/make_golden.py) and, when /root/reference is mounted, against the live reference."""
import os
import sys

import pytest
import torch

from oracle import decoder_oracle as O
from tests.helpers import golden_params, load_golden, rel_l2

REF = "/root/reference"


@pytest.mark.parametrize("name", ["nano", "cfg1", "cfg2"])
def test_forward_and_loss_match_golden(name):
    g = load_golden(name)
    c = g["config"]
    p = golden_params(g)
    with torch.no_grad():
        lg = O.decoder_forward(p, g["tokens"], g["memory"], None, c["H"])
        lgm = O.decoder_forward(p, g["tokens"], g["memory"], g["mem_pad"], c["H"])
    if c["full"]:
        assert (lg - g["logits"]).abs().max() < 2e-5
        assert (lgm - g["logits_mem_pad"]).abs().max() < 2e-5
    else:
        assert (lg[:, :, ::97] - g["logits_sub"]).abs().max() < 5e-5
        assert (lgm[:, :, ::97] - g["logits_mem_pad_sub"]).abs().max() < 5e-5
    loss = O.cross_entropy(lg, g["targets"], 0)
    assert abs(loss.item() - g["loss"]) < 1e-5 * g["loss"] + 1e-6


@pytest.mark.parametrize("name", ["nano", "cfg1", "cfg2"])
def test_gradients_match_golden(name):
    g = load_golden(name)
    c = g["config"]
    p = golden_params(g)
    _, grads = O.loss_and_grads(p, g["tokens"], g["targets"], g["memory"], None, c["H"])
    assert set(grads) == set(g["grad_norm"])
    # fp32 summation order (the restatement's plain matmuls vs torch's fused attention path, 8 generator threads vs
    # this host's) shows more in the 6 x 768 stack at batch 4: up to 1.5e-4 on a gradient norm and 2.2e-3 element-wise
    # on the 512-element subsamples (188 tokens per weight gradient, ReLU units next to zero); the forward agrees to 5e-5
    tol = 1e-4 if c["E"] <= 512 else 5e-4
    tol_sub = 1e-4 if c["E"] <= 512 else 5e-3
    for k, gn in g["grad_norm"].items():
        assert abs(float(grads[k].norm()) - gn) <= tol * gn + 1e-9, k
    if c["full"]:
        for k, v in g["grads"].items():
            assert rel_l2(grads[k], v) < 1e-4, k
    else:
        for k, v in g["grads_sub"].items():
            sub = grads[k].flatten()[::max(1, grads[k].numel() // 512)][:512]
            assert rel_l2(sub, v) < tol_sub, k
    assert float(grads["token_embedding.weight"][0].abs().max()) == 0.0     # padding row


def test_adamw_clip_trajectory_matches_golden():
    g = load_golden("nano")
    c = g["config"]
    p = golden_params(g)
    state, losses = {}, []
    for _ in range(3):
        loss, grads = O.loss_and_grads(p, g["tokens"], g["targets"], g["memory"], None, c["H"])
        O.adamw_step(p, grads, state, lr=g["train_lr"], betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5, max_norm=5.0)
        losses.append(loss.item())
    for a, b in zip(losses, g["train_losses"]):
        assert abs(a - b) < 2e-4 * b
    # Adam moves a coordinate by ~lr per step whatever |g| is, so a coordinate whose true gradient
    # is zero (e.g. the key bias of an attention block: softmax is shift-invariant) follows the SIGN
    # of rounding noise and may differ by up to 2*lr per step; everything else agrees to fp32 accuracy.
    for k, v in g["params_after"].items():
        d = (p[k] - v).abs()
        assert d.max() <= 2 * 3 * g["train_lr"], k
        if "in_proj_bias" not in k:
            assert d.mean() < 5e-6, k


@pytest.mark.parametrize("name", ["nano", "cfg1", "cfg2"])
def test_greedy_matches_golden(name):
    g = load_golden(name)
    c = g["config"]
    p = golden_params(g)
    n = len(g["greedy"])
    with torch.no_grad():
        got = O.greedy_generate(p, g["memory"][:n], 1, 2, g["greedy_max_len"], c["H"])
    assert got == g["greedy"]


def test_beam1_equals_greedy_and_beam_is_sorted():
    g = load_golden("nano")
    c = g["config"]
    p = golden_params(g)
    with torch.no_grad():
        greedy = O.greedy_generate(p, g["memory"][:2], 1, 2, 8, c["H"])
        beam1 = O.beam_generate(p, g["memory"][:2], 1, 2, 8, c["H"], beam_size=1)
        beam3 = O.beam_generate(p, g["memory"][:2], 1, 2, 8, c["H"], beam_size=3)
    assert greedy == beam1
    assert all(s[0] == 1 and len(s) <= 8 for s in beam3)


def test_kv_cache_equivalence_property():
    """Logits at position t from the prefix equal the full-sequence logits at t (post-LN is
    per-position, masks are causal): the property that makes a KV cache exact."""
    g = load_golden("nano")
    c = g["config"]
    p = golden_params(g)
    tok = g["tokens"][:, :9].clone()
    tok[tok == 0] = 5
    with torch.no_grad():
        full = O.decoder_forward(p, tok, g["memory"], None, c["H"])
        for t in (0, 3, 8):
            pre = O.decoder_forward(p, tok[:, :t + 1], g["memory"], None, c["H"])
            assert (pre[:, -1] - full[:, t]).abs().max() < 2e-5


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted (GPU box)")
def test_oracle_matches_live_reference():
    sys.path.insert(0, REF)
    try:
        import config as rcfg
        rcfg.DEVICE = "cpu"
        import decoder as rdec
        V, E, H, L, Fd, ML = 520, 64, 2, 2, 96, 32
        torch.manual_seed(7)
        ref = rdec.TransformerDecoder(V, E, H, L, Fd, ML, dropout=0.0, pad_idx=0).eval()
        p = O.init_params(V, E, H, L, Fd, ML, seed=7)
        sd = ref.state_dict()
        assert set(sd) == set(p) and all(torch.equal(sd[k], p[k]) for k in sd)
        gen = torch.Generator().manual_seed(3)
        tok = torch.randint(4, V, (4, 11), generator=gen)
        tok[:, 0] = 1
        tok[0, 7:] = 0
        tok[2, 4] = 0
        mem = torch.randn(4, 9, E, generator=gen)
        mpm = torch.zeros(4, 9, dtype=torch.bool)
        mpm[1, 5:] = True
        with torch.no_grad():
            assert (ref(tok, mem, None) - O.decoder_forward(p, tok, mem, None, H)).abs().max() < 1e-5
            assert (ref(tok, mem, mpm) - O.decoder_forward(p, tok, mem, mpm, H)).abs().max() < 1e-5
    finally:
        sys.path.remove(REF)
        for m in ("config", "decoder", "utils"):
            sys.modules.pop(m, None)


def test_model_level_oracle_matches_reference_model_golden():
    """tests/golden/model_cfg1.pt comes from the unmodified reference model.ImageToTextModel (BASELINE
    configs[0]); the oracle (CLS token of the frozen tower -> projection -> decoder restatement -> CE)
    must reproduce its logits and loss from the same seed and construction order (model.py:30-114)."""
    from transformers import CLIPConfig, CLIPModel
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_cfg1.pt"), weights_only=True)
    c = g["config"]
    torch.manual_seed(g["seed"])
    tower = CLIPModel(CLIPConfig()).vision_model.eval()           # the reference's (offline) AutoModel.from_pretrained draw
    lin = torch.nn.Linear(768, c["E"])                            # model.py:99
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=None)
    assert abs(float(lin.weight.double().sum()) - g["weight_checksum"]["projection.weight"]) < 1e-6
    assert abs(float(p["fc_out.weight"].double().sum()) - g["weight_checksum"]["decoder.fc_out.weight"]) < 1e-5
    gen = torch.Generator().manual_seed(g["seed"])
    images = torch.randn(c["B"], 3, 224, 224, generator=gen)
    with torch.no_grad():
        cls = tower(pixel_values=images).last_hidden_state[:, 0, :]                      # model.py:141
        mem = O.project_memory(cls.unsqueeze(1), lin.weight, lin.bias)                   # model.py:145-151
        logits = O.decoder_forward(p, g["tokens"], mem, None, c["H"])
    assert abs(float(cls.double().sum()) - g["cls_checksum"]) < 1e-3
    assert (logits[:, :, ::97] - g["logits_sub"]).abs().max() < 5e-5
    loss = O.cross_entropy(logits, g["targets"], 0)
    assert abs(loss.item() - g["loss"]) < 1e-5 * g["loss"]
