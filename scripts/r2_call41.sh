#!/bin/bash
# round 2: BASELINE configs[2]'s size (200 M pairs of 2 x 150 bp, --compare-seq tight) through the drop-in binary on ONE B200;
# the CLI suite on the real engine first (plain inputs now take the discarded-input path by default)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_cli_gpu.py -q -m gpu --timeout 200 -x 2>&1 | tail -4 ) 2>&1 | tail -8
( time timeout 900 python scripts/r2_discard_cli.py --pairs 0 --big-out null > gpurun_out/r02_discard_cli_200M.json 2> gpurun_out/r02_discard_cli_200M.err ) 2>&1 | tail -4
tail -5 gpurun_out/r02_discard_cli_200M.err
cut -c1-900 gpurun_out/r02_discard_cli_200M.json
