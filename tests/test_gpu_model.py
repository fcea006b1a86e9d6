"""Model-level parity on the B200 against the UNMODIFIED reference model.ImageToTextModel
(rows a8 / a9 / a11 of SURVEY section 8): tests/golden/model_cfg1.pt was produced by
tests/golden/make_golden_model.py from /root/reference at BASELINE configs[0] (random-init CLIP ViT-B/32
tower, 768->512 projection, 4-layer d=512 decoder, batch 8, caption length 32, synthetic images).
The weights are rebuilt here from the same seed in the same construction order and verified against
the stored checksums before anything is compared."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_cfg1.pt")


def _build(cuda_dev, g):
    from transformers import CLIPConfig, CLIPImageProcessor, CLIPModel
    from multimodal_image_transformer_b200.model import ImageToTextModel
    c = g["config"]
    torch.manual_seed(g["seed"])
    full = CLIPModel(CLIPConfig())                 # the reference's (patched) AutoModel.from_pretrained draw
    m = ImageToTextModel(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], 0.0, 0, encoder=full.vision_model,
                         encoder_output_dim=768, image_processor=CLIPImageProcessor(), memory_mode="cls", device=cuda_dev)
    sd = m.state_dict()
    for k, v in g["weight_checksum"].items():
        got = float(sd[k].double().sum())
        assert abs(got - v) <= 1e-5 * max(1.0, abs(v)), f"seeded weights differ from the reference's for {k}: {got} vs {v}"
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == g["n_trainable"]
    return m


def _images(g):
    gen = torch.Generator().manual_seed(g["seed"])
    return torch.randn(g["config"]["B"], 3, 224, 224, generator=gen)


def test_model_forward_loss_and_train_steps_vs_reference(cuda_dev):
    from multimodal_image_transformer_b200.train import B200AdamW, fused_train_step
    g = torch.load(GOLDEN, weights_only=True)
    m = _build(cuda_dev, g)
    images, tok, tgt = _images(g).to(cuda_dev), g["tokens"].to(cuda_dev), g["targets"].to(cuda_dev)
    m.eval()
    with torch.no_grad():
        logits = m(images, tok).cpu()
        cls = m.encode(images)[:, 0, :].cpu()
    assert abs(float(cls.double().sum()) - g["cls_checksum"]) < 2e-2 * g["cls_abs_mean"] * cls.numel() ** 0.5   # TF32 conv on the GPU
    err = (logits[:, :, ::97] - g["logits_sub"]).abs().amax(-1) / g["logits_rowmax"]
    assert err.max().item() < 2e-2
    loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), g["targets"].view(-1), ignore_index=0).item()
    assert abs(loss - g["loss"]) < 1e-3 * g["loss"]
    # train.py:80-100 through the fused step (frozen encoder, projection + decoder trained)
    m.train()
    opt = B200AdamW(m, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    losses = [fused_train_step(m, images, tok, tgt, opt, 0, 5.0)[0].item() for _ in range(3)]
    for a, b in zip(losses, g["train_losses"]):
        assert abs(a - b) < 2e-3 * b, (losses, g["train_losses"])
    # and through the reference's own loop: logits -> criterion -> loss.backward() (autograd bridge)
    m2 = _build(cuda_dev, g).train()
    lg = m2(images, tok)
    l2 = torch.nn.functional.cross_entropy(lg.view(-1, lg.shape[-1]), tgt.view(-1), ignore_index=0)
    l2.backward()
    assert abs(l2.item() - g["train_losses"][0]) < 1e-3 * g["train_losses"][0]
    assert m2.projection.weight.grad is not None and torch.isfinite(m2.projection.weight.grad).all()
    assert all(p.grad is None for p in m2.encoder.parameters())


def test_parameter_order_matches_the_reference_for_stock_optimizers(cuda_dev):
    """model.parameters() lists encoder, projection, decoder like the reference (model.py:34-114 registers them in
    that order; train.py:319 builds AdamW over model.parameters()), so a stock torch.optim.AdamW state_dict --
    which maps state by POSITION -- round-trips between the two models."""
    g = torch.load(GOLDEN, weights_only=True)
    m = _build(cuda_dev, g)
    names = [n for n, _ in m.named_parameters()]
    first = [n.split(".")[0] for n in names]
    blocks = [k for i, k in enumerate(first) if i == 0 or first[i - 1] != k]
    assert blocks == ["encoder", "projection", "decoder"], blocks
    dec_names = [n[len("decoder."):] for n in names if n.startswith("decoder.")]
    assert dec_names == [n for n in m.decoder.engine.layout if not n.startswith("projection.")]
    assert dec_names[0] == "token_embedding.weight" and dec_names[-1] == "fc_out.bias"
    trainable = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(trainable, lr=1e-4)
    for p in trainable:
        p.grad = torch.zeros_like(p)
    opt.step()
    sd = opt.state_dict()
    shapes = [tuple(sd["state"][i]["exp_avg"].shape) for i in range(len(trainable))]
    assert shapes[0] == (g["config"]["E"], 768) and shapes[1] == (g["config"]["E"],)          # projection first
    assert shapes[2] == (g["config"]["V"], g["config"]["E"])                                    # then the embedding


def test_model_generate_vs_reference(cuda_dev):
    from PIL import Image
    g = torch.load(GOLDEN, weights_only=True)
    m = _build(cuda_dev, g)
    for i, ref in enumerate(g["generate"]):
        rs = np.random.RandomState(100 + i)
        img = Image.fromarray(rs.randint(0, 256, (224, 224, 3), dtype=np.uint8), "RGB")
        got = m.generate(img, 1, 2, max_len=g["generate_max_len"], method="greedy")
        assert got == ref, (got, ref)


def test_bf16_memory_and_feature_cache(cuda_dev):
    """bf16 features are consumed in place (loss / gradients equal the fp32-input path on bf16-rounded
    features) and FeatureCache runs the frozen tower once per image key."""
    from multimodal_image_transformer_b200.model import FeatureCache
    g = torch.load(GOLDEN, weights_only=True)
    m = _build(cuda_dev, g)
    eng = m.decoder.engine
    images, tok, tgt = _images(g).to(cuda_dev), g["tokens"].to(cuda_dev), g["targets"].to(cuda_dev)
    with torch.no_grad():
        feat32 = m.encode(images)
    feat16 = feat32.to(torch.bfloat16)
    res = []
    for mem in (feat16.float(), feat16):
        eng.zero_grad()
        out = eng.forward_loss(tok, tgt, mem, None, 0, training=True)
        eng.backward()
        res.append((out.cpu().clone(), eng.grads.clone()))
    assert res[0][0][0].item() == res[1][0][0].item()
    assert torch.equal(res[0][1], res[1][1]) or (res[0][1] - res[1][1]).abs().max() < 1e-6 * res[0][1].abs().max() + 1e-9
    # generation accepts bf16 memory too
    eng.decode_begin(feat16, None, beam=1, max_len=8)
    t16, _ = eng.generate_greedy(1, 2, 8)
    eng.decode_begin(feat16.float(), None, beam=1, max_len=8)
    t32, _ = eng.generate_greedy(1, 2, 8)
    assert torch.equal(t16, t32)
    # cache: second epoch never touches the encoder
    calls = {"n": 0}
    orig = m.encoder.forward

    def counted(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    m.encoder.forward = counted
    cache = FeatureCache(m)
    keys = [f"img{i}" for i in range(images.shape[0])]
    a = cache.get(keys, images)
    b = cache.get(keys[::-1], images.flip(0))
    assert calls["n"] == 1 and cache.misses == len(keys) and cache.hits == len(keys)
    assert a.dtype == torch.bfloat16 and torch.equal(a, b.flip(0)) and torch.equal(a, feat16)
    l_cached = m.loss(None, tok, tgt, memory=a, training=False)[0].item()
    l_direct = m.loss(images, tok, tgt, training=False)[0].item()
    assert abs(l_cached - l_direct) < 2e-3 * l_direct


def test_train_one_epoch_and_evaluate_loops_vs_reference(cuda_dev):
    """train.train_one_epoch / train.evaluate (reference train.py:62-151) driven with the reference's batch dicts
    (dataset.collate_fn keys): evaluate returns the golden loss, one epoch of three identical batches returns the mean
    of the golden 3-step trajectory -- with the fused optimizer and with a stock torch.optim.AdamW through logits."""
    from multimodal_image_transformer_b200.train import B200AdamW, evaluate, train_one_epoch
    g = torch.load(GOLDEN, weights_only=True)
    crit = torch.nn.CrossEntropyLoss(ignore_index=0)
    batch = {"images": _images(g), "decoder_input_tokens": g["tokens"], "target_tokens": g["targets"], "image_paths": ["-"] * g["config"]["B"]}
    loader = [batch, batch, batch]
    want = sum(g["train_losses"]) / 3
    m = _build(cuda_dev, g)
    assert abs(evaluate(m, [batch], crit, cuda_dev) - g["loss"]) < 1e-3 * g["loss"]
    opt = B200AdamW(m, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    got = train_one_epoch(m, loader, opt, crit, cuda_dev, 5.0, None, 0, 50, None)
    assert abs(got - want) < 2e-3 * want, (got, want)
    assert m.training                                      # the loop leaves the model in train mode, like the reference
    # the reference's own optimizer type: forward -> logits -> criterion -> backward -> clip -> step
    m2 = _build(cuda_dev, g)
    opt2 = torch.optim.AdamW([p for p in m2.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    got2 = train_one_epoch(m2, loader, opt2, crit, cuda_dev, 5.0, None, 0, 50, None)
    assert abs(got2 - want) < 2e-3 * want, (got2, want)
    # after three steps both optimizers have moved the evaluation loss the same way
    e1, e2 = evaluate(m, [batch], crit, cuda_dev), evaluate(m2, [batch], crit, cuda_dev)
    assert e1 < g["loss"] and abs(e1 - e2) < 2e-3 * e1, (e1, e2, g["loss"])
