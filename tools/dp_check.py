"""N-GPU data-parallel equivalence check (torchrun): the N-rank fused step on per-rank shards must
produce the same loss, gradients and updated parameters as the 1-rank step on the concatenated
batch.  Run: torchrun --nproc-per-node 2 tools/dp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from multimodal_image_transformer_b200.decoder import TransformerDecoder
from multimodal_image_transformer_b200.dp import DataParallel, shard_range
from multimodal_image_transformer_b200.train import B200AdamW, fused_train_step
from tests.helpers import CFGS, synth

rank, local, world = DataParallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
c = dict(CFGS["cfg1"], B=8 * world)
tok, tgt, mem, _ = synth(c, 43)            # padded captions: per-rank non-PAD counts differ


def build():
    torch.manual_seed(42)
    d = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0, device=dev)
    d.train()
    return d, B200AdamW(d, lr=1e-3)

# N-rank run on shards
dec, opt = build()
dp = DataParallel(dec.engine)
dp.broadcast_parameters()
b0, b1 = shard_range(c["B"], rank, world)
out = fused_train_step(dec, mem[b0:b1].to(dev), tok[b0:b1].to(dev), tgt[b0:b1].to(dev), opt, 0, 5.0, dp)
torch.cuda.synchronize()
g_dp = dec.engine.grads.clone()
p_dp = dec.engine.params.clone()
# single-rank run on the whole batch (every rank does it redundantly)
dec1, opt1 = build()
out1 = fused_train_step(dec1, mem.to(dev), tok.to(dev), tgt.to(dev), opt1, 0, 5.0, None)
torch.cuda.synchronize()
g1, p1 = dec1.engine.grads, dec1.engine.params
rel_g = ((g_dp - g1).norm() / g1.norm()).item()
dmax = (p_dp - p1).abs().max().item()
ok = abs(out[0].item() - out1[0].item()) < 1e-4 * out1[0].item() and out[1].item() == out1[1].item() and rel_g < 2e-2 and dmax <= 2e-3
print(f"rank {rank}/{world}: loss dp {out[0].item():.6f} single {out1[0].item():.6f} count {out[1].item()} / {out1[1].item()} "
      f"grad rel_l2 {rel_g:.3e} param max diff {dmax:.2e} {'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
