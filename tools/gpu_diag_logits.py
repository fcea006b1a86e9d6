"""Logit error of the engine against the fp32 oracle next to the noise floor of bf16 storage itself (the bf16-emulating
oracle against the fp32 oracle), per configuration and seed: row-max-relative metric of SURVEY 8c (O1).
usage: python tools/gpu_diag_logits.py cfg2s cfg5s   (B200_ATTN_TC=0 for the mma.sync attention)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import decoder_oracle as O
from tests.helpers import CFGS, make_engine, row_max_rel, synth

dev = torch.device("cuda:0")
print("B200_ATTN_TC =", os.environ.get("B200_ATTN_TC", "(default on)"))
for name in sys.argv[1:] or ["cfg2s", "cfg5s"]:
    c = CFGS[name]
    for seed in (42, 43, 44, 45):
        p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=seed)
        tok, tgt, mem, mpm = synth(c, seed + 1)
        eng = make_engine(c, p, dev)
        with torch.no_grad():
            ref = O.decoder_forward(p, tok, mem, None, c["H"])
            emu = O.decoder_forward(p, tok, mem, None, c["H"], emulate_bf16=True)
        got = eng.forward_logits(tok.to(dev), mem.to(dev), None, training=False)
        print(f"{name} seed {seed}: engine vs fp32 oracle {row_max_rel(got, ref):.4e}   engine vs bf16-emulating oracle "
              f"{row_max_rel(got, emu):.4e}   bf16-emulating vs fp32 oracle {row_max_rel(emu, ref):.4e}")
