export FQD_BENCH_READS=20000000 FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1
for cfg in "2048 5" "1024 5" "4096 5" "2048 4" "2048 3" "3072 5"; do
  set -- $cfg
  out=$(FQD_PP_LEAD=$1 FQD_PP_CTAS=$2 timeout 60 python bench.py --steps 3 --warmup 3 2>/dev/null)
  echo "lead $1 ctas $2 rc $? $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('k1_ms', round(r['avg_launch_ms'],4), 'GBps', round(r['achieved'],1), 'step_ms', round(d['ms_per_step'],3))
except Exception as e: print('no json')")"
done
