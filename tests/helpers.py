"""Shared test utilities: seeded synthetic inputs, golden loading, error metrics."""
import os

import torch

from oracle import decoder_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CFGS = {
    "nano": dict(V=264, E=64, H=2, L=2, F=128, ML=40, B=3, T=17, S=13),
    "tiny": dict(V=1000, E=128, H=2, L=2, F=256, ML=40, B=3, T=17, S=13),
    "hd96": dict(V=2000, E=192, H=2, L=2, F=320, ML=64, B=4, T=47, S=197),
    "cfg1": dict(V=10000, E=512, H=8, L=4, F=2048, ML=100, B=8, T=31, S=50),
    # BASELINE cfg2 decoder at reduced batch (oracle finishes in seconds); full batch in bench.py
    "cfg2s": dict(V=10000, E=768, H=12, L=6, F=3072, ML=100, B=4, T=47, S=197),
    # BASELINE cfg5 decoder (CLIP ViT-L/14 features 257x1024, 12 layers, d=1024, 16 heads) at batch 2
    "cfg5s": dict(V=10000, E=1024, H=16, L=12, F=4096, ML=100, B=2, T=47, S=257),
}


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"decoder_{name}.pt"), weights_only=True)


def golden_params(g):
    c = g["config"]
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=g["seed"])
    for k, v in g["weight_checksum"].items():
        got = float(p[k].double().sum())
        assert abs(got - v) <= 1e-6 * max(1.0, abs(v)), f"seeded weights differ from the reference's for {k}"
    return p


def synth(c, seed=43):
    """tokens: START(1) first, uniform ids in [4,V), PAD(0) tails of random length and one PAD inside a
    prefix; targets likewise; memory ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    B, T, S, V, E = c["B"], c["T"], c["S"], c["V"], c["E"]
    tok = torch.randint(4, V, (B, T), generator=g)
    tok[:, 0] = 1
    tgt = torch.randint(4, V, (B, T), generator=g)
    for b in range(B):
        ln = int(torch.randint(T // 2, T + 1, (1,), generator=g))
        tok[b, ln:] = 0
        tgt[b, max(ln - 1, 1):] = 0
    if B > 1:
        tok[1, 3] = 0
    mem = torch.randn(B, S, E, generator=g)
    mpm = torch.zeros(B, S, dtype=torch.bool)
    mpm[0, S // 2:] = True
    return tok, tgt, mem, mpm


def rel_l2(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def row_max_rel(got, ref):
    """max over rows of max|got-ref| / max|ref| per row: the bf16 logits metric of SURVEY 8c (O1)."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    return ((got - ref).abs().amax(-1) / ref.abs().amax(-1).clamp_min(1e-20)).max().item()


def make_engine(c, params, device, enc_dim=None):
    from multimodal_image_transformer_b200.engine import DecoderEngine
    eng = DecoderEngine(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], pad_idx=0, enc_dim=enc_dim, device=device)
    eng.load(params)
    return eng
