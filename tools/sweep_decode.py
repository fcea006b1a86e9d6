"""Sweeps the generation tuning knobs (read per decode plan) in ONE process at BASELINE configs[3] shapes and
prints one JSON line per setting: ms per 512-caption batch (min over graph replays), ms per position, and the
fraction of token ids identical to the default setting's (hints must not change results).

  python tools/sweep_decode.py [--beam 4] [--quick] [--l2] [--fat1] [--stream] [--deep] [--dyn]

Without a mode flag only the split-K / unpaired-GEMM settings are swept; --l2: L2 warm-up and eviction hints, --fat1: the
fat-CTA attention on one partition, --stream: partitions x SM budget x dedicated attention stream, --deep: 3-deep rings,
--dyn: dynamic item scheduling and weights evict-last (the four sweeps of profiles/r01_decode_sweeps.txt).
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_image_transformer_b200.engine import DecoderEngine

KNOBS = ("B200_DEC_PREFETCH_MB", "B200_DEC_KV_FLAGS", "B200_DECODE_PARTS", "B200_DEC_KSPLIT_E", "B200_DEC_KSPLIT_F",
         "B200_DEC_SINGLE_CTA", "B200_DEC_ATTN_GRID", "B200_DEC_GEMM_CTAS", "B200_DEC_ATTN_STREAM", "B200_DEC_ATTN_DYN")


def main():
    beam = int(sys.argv[sys.argv.index("--beam") + 1]) if "--beam" in sys.argv else 1
    quick = "--quick" in sys.argv
    dev = torch.device("cuda:0")
    V, E, H, L, F, ML = 10000, 768, 12, 6, 3072, 100
    B, S, max_len = 512, 197, 48
    eng = DecoderEngine(V, E, H, L, F, ML, device=dev)
    torch.manual_seed(0)
    eng.params.normal_(0, 0.02)
    eng.sync_shadow(force=True)
    mem = torch.randn(B, S, E, device=dev)

    def run(cfg, reps=5):
        for k in KNOBS:
            os.environ.pop(k, None)
        for k, v in cfg.items():
            os.environ[k] = str(v)
        eng.decode_begin(mem, None, beam=beam, max_len=max_len)
        best, toks = 1e9, None
        for it in range(reps):
            toks = None     # release the previous output first: the beam graph is keyed on the output address
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if beam == 1:
                toks, _ = eng.generate_greedy(1, V + 7, max_len, 0)
            else:
                toks = eng.generate_beam(1, V + 7, max_len)[0]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= 2:
                best = min(best, dt)
        return best * 1e3, toks.clone()

    base_ms, base_toks = run({})
    print(json.dumps({"cfg": {}, "ms": round(base_ms, 3), "ms_per_pos": round(base_ms / (max_len - 1), 4)}), flush=True)
    cfgs = []
    if "--l2" in sys.argv:      # round-1 sweep of the L2 warm-up / eviction hints (measured: no gain)
        pf_opts = (32, 64, 96, 128) if quick else (16, 32, 48, 64, 80, 96, 112, 128, 160, 200)
        for fl in (0, 1, 2, 3):
            for pf in (0,) + pf_opts:
                if pf == 0 and fl == 0:
                    continue
                cfgs.append({"B200_DEC_PREFETCH_MB": pf, "B200_DEC_KV_FLAGS": fl})
    small = {"B200_DEC_SINGLE_CTA": 1, "B200_DEC_KSPLIT_E": 2, "B200_DEC_KSPLIT_F": 4, "B200_DEC_KV_FLAGS": 3}
    cfgs += [{"B200_DEC_SINGLE_CTA": 1}, {"B200_DEC_KSPLIT_E": 2, "B200_DEC_KSPLIT_F": 4}, dict(small)]
    cfgs += [dict(small, B200_DEC_KSPLIT_E=ke, B200_DEC_KSPLIT_F=kf) for ke, kf in ((3, 4), (2, 3), (1, 4), (2, 6))]
    if "--fat1" in sys.argv:
        # the fat-CTA attention alone (one partition): cost of the restructuring (GEMMs uncapped)
        cfgs += [dict(small, B200_DEC_ATTN_GRID=g, B200_DEC_GEMM_CTAS=148) for g in (148, 128, 112, 96)]
    if "--stream" in sys.argv:
        # concurrent partitions: SM budget of the attention stream, with / without the dedicated attention stream
        for parts in ((4,) if quick else (3, 4, 5, 6, 8)):
            for g in ((104,) if quick else (96, 104, 112, 120)):
                for ast in (0, 1):
                    cfgs.append(dict(small, B200_DECODE_PARTS=parts, B200_DEC_ATTN_GRID=g, B200_DEC_ATTN_STREAM=ast))
    if "--deep" in sys.argv:       # 3-deep rings in the fat-CTA attention (flags bit 3), SM budget, partitions
        for parts in (4, 3, 2):
            for g in (88, 96, 104, 112, 120):
                for fl in (3, 11):
                    cfgs.append({"B200_DECODE_PARTS": parts, "B200_DEC_ATTN_GRID": g, "B200_DEC_KV_FLAGS": fl})
    if "--dyn" in sys.argv:
        # dynamic item scheduling of the cross attention, weights evict-last (flags bit 2)
        best = dict(small, B200_DECODE_PARTS=4, B200_DEC_ATTN_GRID=104)
        cfgs += [dict(best), dict(best, B200_DEC_ATTN_DYN=1), dict(best, B200_DEC_KV_FLAGS=7), dict(best, B200_DEC_ATTN_DYN=1, B200_DEC_KV_FLAGS=7),
                 dict(best, B200_DEC_ATTN_DYN=1, B200_DEC_KV_FLAGS=6), dict(best, B200_DEC_ATTN_DYN=1, B200_DEC_KV_FLAGS=0)]
        cfgs += [dict(small, B200_DEC_ATTN_DYN=1), dict(small, B200_DEC_KV_FLAGS=7), dict(small, B200_DEC_ATTN_DYN=1, B200_DEC_KV_FLAGS=7)]
        if not quick:
            for parts in (2, 3, 4):
                for g in (0, 88, 96, 104, 112, 120, 128):
                    c = dict(small, B200_DECODE_PARTS=parts, B200_DEC_ATTN_DYN=1)
                    if g:
                        c["B200_DEC_ATTN_GRID"] = g
                    cfgs.append(c)
                    cfgs.append(dict(c, B200_DEC_KV_FLAGS=7))
            cfgs += [dict(best, B200_DEC_ATTN_DYN=1, B200_DEC_ATTN_STREAM=1), dict(best, B200_DEC_ATTN_DYN=1, B200_DEC_ATTN_STREAM=1, B200_DEC_KV_FLAGS=7)]
    results = []
    for cfg in cfgs:
        try:
            ms, toks = run(cfg)
            same = float((toks == base_toks).float().mean())
            rec = {"cfg": cfg, "ms": round(ms, 3), "ms_per_pos": round(ms / (max_len - 1), 4), "vs_default": round(base_ms / ms, 4),
                   "ids_equal": round(same, 5)}
        except Exception as ex:  # keep sweeping: a failing knob combination is a result too
            rec = {"cfg": cfg, "error": str(ex)[:300]}
        results.append(rec)
        print(json.dumps(rec), flush=True)
    ok = [r for r in results if "ms" in r]
    ok.sort(key=lambda r: r["ms"])
    print("BEST", json.dumps(ok[:5]), flush=True)
    # the default again at the end (drift check)
    ms, _ = run({})
    print(json.dumps({"cfg": "default-again", "ms": round(ms, 3)}), flush=True)


if __name__ == "__main__":
    main()
