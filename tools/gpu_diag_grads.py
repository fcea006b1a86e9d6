"""Per-tensor gradient error of the engine vs the fp32 oracle and the bf16-emulating oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import decoder_oracle as O
from tests.helpers import CFGS, make_engine, rel_l2, synth

dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["nano", "cfg1"]:
    c = CFGS[name]
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=42)
    tok, tgt, mem, _ = synth(c, 43)
    eng = make_engine(c, p, dev)
    _, g32 = O.loss_and_grads(p, tok, tgt, mem, None, c["H"])
    _, g16 = O.loss_and_grads(p, tok, tgt, mem, None, c["H"], emulate_bf16=True)
    eng.zero_grad()
    eng.forward_loss(tok.to(dev), tgt.to(dev), mem.to(dev), None, 0, training=True)
    eng.backward()
    torch.cuda.synchronize()
    print(f"== {name}: tensor, rel_l2 vs emulated-bf16 oracle, vs fp32 oracle, oracle16-vs-oracle32, cos(emul)")
    for k in g32:
        got = eng.view(k, eng.grads).float().cpu()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), g16[k].flatten(), dim=0).item()
        print(f"  {k:62s} {rel_l2(got, g16[k]):.3e} {rel_l2(got, g32[k]):.3e} {rel_l2(g16[k], g32[k]):.3e} {cos:.5f}")
