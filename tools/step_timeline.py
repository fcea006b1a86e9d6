"""In-situ kernel timeline of the graph-replayed train step (torch.profiler / CUPTI): per-kernel time inside the replay
(warm caches, concurrent side streams) and the idle gaps on the main stream -- the numbers the cold-cache, serialised ncu
launch list cannot give.  usage: python tools/step_timeline.py [cfg2|cfg5]"""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_image_transformer_b200.decoder import TransformerDecoder
from multimodal_image_transformer_b200.train import B200AdamW, GraphedTrainStep

from multimodal_image_transformer_b200.dp import DataParallel
os.environ.setdefault("NCCL_DEBUG", "NONE")
c = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
rank, local, world = DataParallel.init_from_env("nccl")       # under torchrun: the data-parallel step, rank 0 reports
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.manual_seed(42)
dec = TransformerDecoder(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], dropout=0.0, pad_idx=0, device=dev)
dec.train()
opt = B200AdamW(dec, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5)
dp = DataParallel(dec.engine) if world > 1 else None
if dp is not None:
    dp.broadcast_parameters()
tok, tgt, mem = bench.synth_batch(c, 1000 + rank)
tok, tgt, mem = tok.to(dev), tgt.to(dev), mem.to(dev, torch.bfloat16)
gs = GraphedTrainStep(dec, opt, 0, 5.0, warmup=2, dp=dp)
for _ in range(6):
    gs(mem, tok, tgt)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        gs(mem, tok, tgt)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.cuda_time_total is not None]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()), key=lambda x: x[0])
if rank != 0:
    sys.exit(0)
if not ks:
    print("no kernel records (CUPTI unavailable?)"); sys.exit(0)
# split into replays by the adamw kernel
ends = [i for i, k in enumerate(ks) if "adamw_kernel" in k[2]]
a, b = ends[0] + 1, ends[1] + 1
step = ks[a:b]
t0, t1 = step[0][0], max(k[1] for k in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in step:
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("b200::", "")
    agg[n][0] += 1; agg[n][1] += e - s
busy = 0.0; cur_end = t0
for s, e, n in sorted(step):
    if e > cur_end:
        busy += e - max(s, cur_end); cur_end = e
print(f"step span {t1 - t0:.1f} us, union of kernel intervals {busy:.1f} us, idle {t1 - t0 - busy:.1f} us, {len(step)} kernels")
if len(ends) > 2:
    nxt = ks[b:ends[2] + 1]
    print(f"gap between this replay's last kernel and the next replay's first kernel: {nxt[0][0] - t1:.1f} us; "
          f"replay period {nxt[0][0] - t0:.1f} us")
    first = sorted(step)[:4]
    print("first kernels:", [(round(s_ - t0, 1), round(e_ - t0, 1), n_[:30]) for s_, e_, n_ in first])
    last = sorted(step, key=lambda k: k[1])[-3:]
    print("last kernels:", [(round(s_ - t0, 1), round(e_ - t0, 1), n_[:30]) for s_, e_, n_ in last])
for n, (cnt, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.1f} us {cnt:4d} {t / cnt:8.1f} us/launch  {n[:80]}")
nccl = sorted((s_ - t0, e_ - t0) for s_, e_, n_ in step if "nccl" in n_.lower())
if nccl:
    last_compute_before_opt = max(e_ for s_, e_, n_ in step if "gemm" in n_ or "embed_bwd" in n_) - t0
    print("NCCL kernels (start, end) us:", [(round(a_), round(b_)) for a_, b_ in nccl])
    print(f"last backward kernel ends at {last_compute_before_opt:.0f} us; last NCCL kernel ends at {nccl[-1][1]:.0f} us; step ends at {t1 - t0:.0f} us")
