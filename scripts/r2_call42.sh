#!/bin/bash
# round 2: BASELINE configs[2]'s size (200 M pairs of 2 x 150 bp, --compare-seq tight) through the drop-in binary on ONE B200
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python scripts/r2_discard_cli.py --pairs 0 --big-out null > gpurun_out/r02_discard_cli_200M.json 2> gpurun_out/r02_discard_cli_200M.err ) 2>&1 | tail -4
tail -5 gpurun_out/r02_discard_cli_200M.err
cut -c1-900 gpurun_out/r02_discard_cli_200M.json
