#!/bin/bash
# A/B of engine switches on ONE box: each line is a set of env assignments; prints graph-replay ms/step.
# usage: bash tools/ab_bench.sh "A=1" "B=1 C=2" ...   (an empty string "" = defaults)
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-decode 2>/dev/null | tail -1)
  python - "$cfg" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(f"[{sys.argv[1]:45s}] graph {d['ms_per_step']:.3f} ms  eager {d.get('ms_per_step_eager', 0):.3f} ms  e2e {d['e2e']['ms_per_step']:.3f} ms  "
      f"gemm {d['roofline']['gemm_ms_per_step']:.3f} ms ({d['roofline']['frac']:.3f})  launches/step {d['gpu_launches']/d['steps']:.0f}  sm {d['clocks']['sm_mhz']}")
PY
done
