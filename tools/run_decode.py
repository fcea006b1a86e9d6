"""Runs KV-cached greedy generation once at BASELINE configs[3] shapes (for ncu captures / timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200.engine import DecoderEngine
dev = torch.device("cuda:0")
V, E, H, L, F, ML = 10000, 768, 12, 6, 3072, 100
B, S, max_len = 512, 197, int(os.environ.get("MAXLEN", "48"))
eng = DecoderEngine(V, E, H, L, F, ML, device=dev)
torch.manual_seed(0)
eng.params.normal_(0, 0.02); eng.sync_shadow(force=True)
mem = torch.randn(B, S, E, device=dev)
for it in range(int(os.environ.get("REPS", "3"))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.decode_begin(mem, None, beam=1, max_len=max_len)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    toks, lens = eng.generate_greedy(1, V + 7, max_len, 0)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {it}: begin {1e3*(t1-t0):.2f} ms, generate {1e3*(t2-t1):.2f} ms ({1e3*(t2-t1)/(max_len-1):.3f} ms/step)")
