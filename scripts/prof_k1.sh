# usage: scripts/prof_k1.sh <tag>   - lead sweep + ncu full capture of k_parse_pack (run under gpurun)
tag=${1:-k1}
export FQD_BENCH_READS=20000000 FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1
for lead in ${LEADS:-768 1024 1536 2048}; do
  FQD_PP_LEAD=$lead timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('lead $lead', 'k1_ms', round(r['avg_launch_ms'],4), 'GBps', round(r['achieved'],1), 'step_ms', round(d['ms_per_step'],3))"
done
export FQD_BENCH_READS=6000000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_parse_pack -s 3 -c 1 -f -o gpurun_out/prof_parse_$tag python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
