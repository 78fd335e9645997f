#!/bin/bash
cd "$(dirname "$0")/.."
( time timeout 1400 python -m pytest tests -x -q -m gpu --timeout 200 --durations=25 2>&1 | tail -45 ) 2>&1
