"""Generates tests/golden/model_cfg1.pt by running the UNMODIFIED reference model.ImageToTextModel
(/root/reference/model.py: forward :116-169, generate :171-255) at BASELINE configs[0]: a
random-init CLIP ViT-B/32 vision tower (50 tokens x 768, frozen) + 768->512 projection + 4-layer
d=512 8-head decoder, batch 8, caption length 32, synthetic 224x224 images.  Run in the build
container only:

    python tests/golden/make_golden_model.py

There is no network, so `from_pretrained` is pointed at a seeded random initialisation
(CLIPModel(CLIPConfig()) / CLIPImageProcessor()); nothing else of the reference is touched.  Weights
are not stored: the GPU test rebuilds them from the same seed in the same construction order
(encoder, projection, decoder) and checks the stored checksums first.
"""
import os
import sys

import numpy as np
import torch
from PIL import Image
from transformers import CLIPConfig, CLIPImageProcessor, CLIPModel

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import config  # noqa: E402  (reference)
config.DEVICE = "cpu"
config.ENCODER_MODEL_NAME = "openai/clip-vit-base-patch32"
config.IMAGE_PROCESSOR_NAME = "openai/clip-vit-base-patch32"
import model as refmodel  # noqa: E402  (reference)

refmodel.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: CLIPModel(CLIPConfig()))
refmodel.AutoImageProcessor.from_pretrained = staticmethod(lambda name, *a, **k: CLIPImageProcessor())

C = dict(V=10000, E=512, H=8, L=4, F=2048, ML=100, B=8, T=31)
SEED = 42


def synth_tokens(seed):
    g = torch.Generator().manual_seed(seed)
    B, T, V = C["B"], C["T"], C["V"]
    tok = torch.randint(4, V, (B, T), generator=g)
    tok[:, 0] = 1
    tgt = torch.randint(4, V, (B, T), generator=g)
    for b in range(B):
        ln = int(torch.randint(T // 2, T + 1, (1,), generator=g))
        tok[b, ln:] = 0
        tgt[b, max(ln - 1, 1):] = 0
    return tok, tgt


def test_image(i):
    rs = np.random.RandomState(100 + i)
    return Image.fromarray(rs.randint(0, 256, (224, 224, 3), dtype=np.uint8), "RGB")


def main():
    torch.set_num_threads(8)
    torch.manual_seed(SEED)
    m = refmodel.ImageToTextModel(C["V"], C["E"], C["H"], C["L"], C["F"], C["ML"], 0.0, 0)
    g = torch.Generator().manual_seed(SEED)
    images = torch.randn(C["B"], 3, 224, 224, generator=g)
    tok, tgt = synth_tokens(SEED + 1)
    out = {"config": dict(C), "seed": SEED, "tokens": tok, "targets": tgt, "torch_version": str(torch.__version__)}
    sd = m.state_dict()
    picks = ["projection.weight", "projection.bias", "decoder.token_embedding.weight", "decoder.fc_out.weight",
             "decoder.transformer_decoder.layers.3.linear2.weight",
             "encoder.embeddings.patch_embedding.weight", "encoder.encoder.layers.11.mlp.fc2.weight"]
    out["weight_checksum"] = {k: float(sd[k].double().sum()) for k in picks}
    out["n_trainable"] = sum(p.numel() for p in m.parameters() if p.requires_grad)
    m.eval()
    with torch.no_grad():
        logits = m(images, tok)
        cls = m.encoder(pixel_values=images).last_hidden_state[:, 0, :]
    out["cls_checksum"] = float(cls.double().sum())
    out["cls_abs_mean"] = float(cls.abs().mean())
    crit = torch.nn.CrossEntropyLoss(ignore_index=0)
    out["loss"] = float(crit(logits.view(-1, C["V"]), tgt.reshape(-1)))
    out["logits_sub"] = logits[:, :, ::97].clone()
    out["logits_rowmax"] = logits.abs().amax(-1)
    # three optimisation steps exactly as train.py:80-100 (AdamW over model.parameters(), clip 5.0)
    m.train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
    traj = []
    for _ in range(3):
        opt.zero_grad()
        loss = crit(m(images, tok).view(-1, C["V"]), tgt.reshape(-1))
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
        opt.step()
        traj.append(float(loss))
    out["train_losses"] = traj
    # generate() of the reference on the INITIAL weights (rebuild from the seed), two PIL images
    torch.manual_seed(SEED)
    m = refmodel.ImageToTextModel(C["V"], C["E"], C["H"], C["L"], C["F"], C["ML"], 0.0, 0).eval()
    out["generate_max_len"] = 12
    out["generate"] = [m.generate(test_image(i), 1, 2, max_len=12, method="greedy") for i in range(2)]
    path = os.path.join(HERE, "model_cfg1.pt")
    torch.save(out, path)
    print("loss", out["loss"], "traj", traj, "generate", out["generate"], "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
