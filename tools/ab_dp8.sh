#!/bin/bash
# A/B of data-parallel settings on ONE 8-GPU box: each argument is a set of env assignments
for cfg in "$@"; do
  out=$(env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-decode --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$cfg" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(f"[{sys.argv[1]:40s}] graph {d['ms_per_step']:.3f} ms  eager {d.get('ms_per_step_eager', 0):.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  gemm {d['roofline']['gemm_ms_per_step']:.3f} ms ({d['roofline']['frac']:.3f})  {d['value']/1e6:.3f} Mtok/s")
PY
done
