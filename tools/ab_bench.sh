#!/usr/bin/env bash
# same-box A/B of the train step: tools/ab_bench.sh "ENV=1 ..." "ENV=2 ..." (each argument = one arm's environment)
for arm in "$@"; do
  echo "== $arm"
  env $arm python bench.py --no-cpu-baseline --no-decode --no-varlen --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f  eager %.3f  gemm_ms %.3f  gemm TF/s %.0f  clocks %s' % (d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['gemm_ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz']))"
done
