#!/bin/bash
# round 2, discarded-input path (inputs larger than device memory): parity tests, then the binary at size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
free -g | sed -n 2p; nproc; df -h /dev/shm | tail -1
( time timeout 600 python -m pytest tests/test_seq_gpu.py tests/test_unordered_gpu.py tests/test_differential_gpu.py -q -m gpu -k "discard" --timeout 200 -x 2>&1 | tail -6 ) 2>&1 | tail -10
( time timeout 900 python scripts/r2_discard_cli.py --pairs 25000000 > gpurun_out/r02_discard_cli.json 2> gpurun_out/r02_discard_cli.err ) 2>&1 | tail -4
tail -5 gpurun_out/r02_discard_cli.err
cut -c1-700 gpurun_out/r02_discard_cli.json
