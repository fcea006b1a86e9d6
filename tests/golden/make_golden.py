"""Generates tests/golden/*.pt by running the UNMODIFIED reference from /root/reference
(decoder.TransformerDecoder, nn.CrossEntropyLoss, clip_grad_norm_, torch.optim.AdamW and the greedy
loop of model.py:216-242) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py [case ...]        (default: every case)

The reference cannot travel to the GPU box, the fixtures do.  Weights are NOT stored: they are
reproduced from the seed by oracle.decoder_oracle.init_params, which constructs the same torch.nn
modules in the same order as decoder.py:105-132 (bit-identical for one torch build); a checksum of
the weights is stored so that a mismatch is detected rather than silently tolerated.
"""
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import config  # noqa: E402  (reference)
config.DEVICE = "cpu"   # utils.py:70 moves masks to the global device
import decoder as refdec  # noqa: E402  (reference)

CASES = {
    # name: V, E, H, L, F, max_len, B, T, S, full (store everything) or subsampled
    "nano": dict(V=264, E=64, H=2, L=2, F=128, ML=40, B=3, T=17, S=13, full=True),
    "cfg1": dict(V=10000, E=512, H=8, L=4, F=2048, ML=100, B=8, T=31, S=50, full=False),
    # BASELINE configs[1] decoder (the headline shape: ViT-B/16 features 197 x 768, 6 layers, d = 768, 12 heads) at batch 4
    "cfg2": dict(V=10000, E=768, H=12, L=6, F=3072, ML=100, B=4, T=47, S=197, full=False),
}


def synth(c, seed):
    g = torch.Generator().manual_seed(seed)
    B, T, S, V, E = c["B"], c["T"], c["S"], c["V"], c["E"]
    tok = torch.randint(4, V, (B, T), generator=g)
    tok[:, 0] = 1                                   # START (config.py:117)
    tgt = torch.randint(4, V, (B, T), generator=g)
    for b in range(B):
        ln = int(torch.randint(T // 2, T + 1, (1,), generator=g))
        tok[b, ln:] = 0
        tgt[b, max(ln - 1, 1):] = 0
    tok[1, 3] = 0                                   # a PAD inside the prefix
    mem = torch.randn(B, S, E, generator=g)
    mpm = torch.zeros(B, S, dtype=torch.bool)
    mpm[0, S // 2:] = True
    return tok, tgt, mem, mpm


def main():
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or list(CASES)):
        c = CASES[name]
        seed = 42
        torch.manual_seed(seed)
        model = refdec.TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0)
        tok, tgt, mem, mpm = synth(c, seed + 1)
        out = {"config": dict(c), "seed": seed, "tokens": tok, "targets": tgt, "memory": mem, "mem_pad": mpm,
               "torch_version": str(torch.__version__)}
        sd = model.state_dict()
        out["weight_checksum"] = {k: float(v.double().sum()) for k, v in sd.items()}
        model.eval()
        with torch.no_grad():
            logits = model(tok, mem, None)
            logits_m = model(tok, mem, mpm)
        crit = torch.nn.CrossEntropyLoss(ignore_index=0)
        out["loss"] = float(crit(logits.view(-1, c["V"]), tgt.reshape(-1)))
        if c["full"]:
            out["logits"] = logits.clone()
            out["logits_mem_pad"] = logits_m.clone()
        else:
            out["logits_sub"] = logits[:, :, ::97].clone()
            out["logits_mem_pad_sub"] = logits_m[:, :, ::97].clone()
            out["logits_rowmax"] = logits.abs().amax(-1)
        # gradients (train.py:83-93)
        model.train()
        model.zero_grad()
        loss = crit(model(tok, mem, None).view(-1, c["V"]), tgt.reshape(-1))
        loss.backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        out["grad_norm"] = {k: float(v.norm()) for k, v in grads.items()}
        if c["full"]:
            out["grads"] = grads
        else:
            out["grads_sub"] = {k: v.flatten()[::max(1, v.numel() // 512)][:512].clone() for k, v in grads.items()}
        # 3 optimizer steps exactly as train.py:80-100 with config.py:80-90 hyper-parameters
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3 if c["full"] else 1e-4, betas=(0.9, 0.98), eps=1e-9,
                                weight_decay=1e-5)
        traj = []
        for _ in range(3):
            opt.zero_grad()
            loss = crit(model(tok, mem, None).view(-1, c["V"]), tgt.reshape(-1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()
            traj.append(float(loss))
        out["train_lr"] = 1e-3 if c["full"] else 1e-4
        out["train_losses"] = traj
        out["param_norm_after"] = {k: float(p.detach().norm()) for k, p in model.named_parameters()}
        if c["full"]:
            out["params_after"] = {k: p.detach().clone() for k, p in model.named_parameters()}
        # greedy generation with the reference loop (model.py:216-242) on the INITIAL weights
        torch.manual_seed(seed)
        model = refdec.TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0).eval()
        gen = []
        steps = 12
        with torch.no_grad():
            for b in range(min(c["B"], 4)):
                ids = torch.tensor([[1]], dtype=torch.long)
                for _ in range(steps - 1):
                    lg = model(tgt_tokens=ids, memory=mem[b:b + 1], memory_padding_mask=None)
                    nxt = torch.argmax(lg[:, -1, :], dim=-1).unsqueeze(0)
                    ids = torch.cat([ids, nxt], dim=1)
                    if nxt.item() == 2:
                        break
                gen.append(ids[0].tolist())
        out["greedy_max_len"] = steps
        out["greedy"] = gen
        path = os.path.join(HERE, f"decoder_{name}.pt")
        torch.save(out, path)
        print(name, "loss", out["loss"], "traj", traj, "greedy", gen[0][:6], "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
