for d in 0.0 0.1; do python bench.py --dropout $d --no-cpu-baseline --no-decode --no-varlen 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dropout', d['config']['dropout'], 'ms/step', round(d['ms_per_step'],3), 'eager', round(d['ms_per_step_eager'],3), 'loss', d['last_loss'])"; done
B200_ATTN_TC_BWD=0 python bench.py --dropout 0.1 --no-cpu-baseline --no-decode --no-varlen 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tc bwd off: dropout', d['config']['dropout'], 'ms/step', round(d['ms_per_step'],3))"
