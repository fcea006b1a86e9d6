run() { echo "$@"; env "$@" timeout 60 python tools/gpu_diag_attn_tc.py --bench cross 2>&1 | tail -2; }
run A=0
run B200_ATTN_TC_WAIT=1
run B200_ATTN_TC_WAIT=2
