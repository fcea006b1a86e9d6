"""In-situ kernel timeline (torch.profiler / CUPTI) of the graph-replayed greedy generation at BASELINE configs[3]:
per-kernel time inside the replay (four concurrent partitions, warm caches) and the union / idle time."""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200.engine import DecoderEngine
dev = torch.device("cuda:0")
V, E, H, L, F, ML = 10000, 768, 12, 6, 3072, 100
B, S, max_len = 512, 197, 48
eng = DecoderEngine(V, E, H, L, F, ML, device=dev)
torch.manual_seed(0)
eng.params.normal_(0, 0.02); eng.sync_shadow(force=True)
mem = torch.randn(B, S, E, device=dev)
for _ in range(3):
    eng.decode_begin(mem, None, beam=1, max_len=max_len)
    eng.generate_greedy(1, V + 7, max_len, 0)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.decode_begin(mem, None, beam=1, max_len=max_len)
    eng.generate_greedy(1, V + 7, max_len, 0)
    torch.cuda.synchronize()
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()),
            key=lambda x: x[0])
t0, t1 = ks[0][0], max(k[1] for k in ks)
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("b200::", "")
    agg[n][0] += 1; agg[n][1] += e - s
busy, cur = 0.0, t0
for s, e, n in ks:
    if e > cur:
        busy += e - max(s, cur); cur = e
print(f"span {(t1 - t0) / 1e3:.2f} ms, union {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms, {len(ks)} kernels")
tot = sum(v[1] for v in agg.values())
for n, (cnt, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{t / 1e3:8.2f} ms {100 * t / tot:5.1f}% {cnt:6d} {t / cnt:7.1f} us/launch  {n[:70]}")
