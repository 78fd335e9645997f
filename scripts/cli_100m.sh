# BASELINE configs[1] (100 M x 150 bp SE, --fast) through the drop-in binary, files in / files out on tmpfs.
avail_kb=$(awk '/MemAvailable/ {print $2}' /proc/meminfo); shm_kb=$(df -k /dev/shm | awk 'NR==2 {print $4}')
echo "MemAvailable ${avail_kb} kB, /dev/shm free ${shm_kb} kB"
reads=100000000
# input 322 B + output ~226 B per read on tmpfs, plus page cache head room
need_kb=$((reads / 1000 * 322 * 2))
if [ "$avail_kb" -lt $((need_kb * 2)) ] || [ "$shm_kb" -lt $((need_kb + need_kb / 4)) ]; then reads=40000000; fi
echo "reads=$reads"
timeout 200 python bench_cli.py --reads $reads --ref-reads 2000000 --formats plain --repeats 1 > gpurun_out/bench_cli_100M.json 2> gpurun_out/bench_cli_100M.err; tail -3 gpurun_out/bench_cli_100M.err; cut -c1-330 gpurun_out/bench_cli_100M.json
