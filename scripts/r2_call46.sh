#!/bin/bash
# last check of the round: the binary's spooled inputs (.gz, FIFO) on the real engine after the spool moved behind an asynchronous writer
cd "$(dirname "$0")/.."
( time timeout 110 python -m pytest tests/test_differential_gpu.py tests/test_cli_gpu.py -q -m gpu --timeout 60 -x -k "sequence_modes_discarded_input or unordered_discarded_input or fifo or pipe_input" 2>&1 | tail -4 ) 2>&1 | tail -8
