"""Why does bench.py's graph-replay loop take longer per step than one replay's kernel span?  Times 20 replays
(CUDA events) of the same captured train step: one graph, two alternating graphs (bench.py's loop), and each with the NVML
clock sampler thread of bench.py running."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_image_transformer_b200.decoder import TransformerDecoder
from multimodal_image_transformer_b200.train import B200AdamW, GraphedTrainStep

c = bench.CONFIGS["cfg2"]
dev = torch.device("cuda:0")
torch.manual_seed(42)
dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0, device=dev)
dec.train()
opt = B200AdamW(dec, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
batches = []
for i in range(2):
    tok, tgt, mem = bench.synth_batch(c, 1000 + i)
    batches.append((tok.to(dev), tgt.to(dev), mem.to(dev, torch.bfloat16)))
gs = [GraphedTrainStep(dec, opt, 0, 5.0, warmup=0) for _ in range(2)]
for g, (tok, tgt, mem) in zip(gs, batches):
    for _ in range(3):
        g(mem, tok, tgt)
torch.cuda.synchronize()

def timed(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

one = lambda i: gs[0](batches[0][2], batches[0][0], batches[0][1])
two = lambda i: gs[i % 2](batches[i % 2][2], batches[i % 2][0], batches[i % 2][1])
raw = lambda i: gs[0].graph.replay()
print(f"one graph            : {timed(one):.3f} ms/step")
print(f"graph.replay() only  : {timed(raw):.3f} ms/step")
print(f"two alternating      : {timed(two):.3f} ms/step")
s = bench.ClockSampler(0)
s.start()
print(f"one graph + sampler  : {timed(one):.3f} ms/step")
print(f"two graphs + sampler : {timed(two):.3f} ms/step")
print("clocks:", s.stop())
print(f"one graph (again)    : {timed(one):.3f} ms/step")
