# Where does the drop-in binary spend its wall clock?  FQD_TRACE=1 checkpoints for 0.1 M / 2 M / 20 M reads.
python - <<'PY'
import bench_cli, pathlib
p = pathlib.Path("/dev/shm/t20.fq"); bench_cli.synth_file(p, 20_000_000)
d = p.read_bytes()
pathlib.Path("/dev/shm/t2.fq").write_bytes(d[:2_000_000 * 322]); pathlib.Path("/dev/shm/t01.fq").write_bytes(d[:100_000 * 322])
PY
E=fastq-dupaway_b200/host/fastq-dupaway
for th in auto 1; do for f in t01 t2 t20 t20; do
  echo "== $f threads=$th"; 
  if [ $th = auto ]; then unset FQD_IO_THREADS; else export FQD_IO_THREADS=$th; fi
  s=$(date +%s.%N); FQD_TRACE=1 $E -i /dev/shm/$f.fq -o /dev/shm/out.fq --fast 2>&1 | grep -v "^\[trace\]" | tail -12; e=$(date +%s.%N); echo "wall $(echo "$e - $s" | bc)"
done; done
echo "== block 32 MiB, threads auto"; unset FQD_IO_THREADS
s=$(date +%s.%N); FQD_BLOCK_BYTES=33554432 FQD_TRACE=1 $E -i /dev/shm/t20.fq -o /dev/shm/out.fq --fast 2>&1 | tail -12; e=$(date +%s.%N); echo "wall $(echo "$e - $s" | bc)"
rm -f /dev/shm/t*.fq /dev/shm/out.fq
