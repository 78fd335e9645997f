#!/usr/bin/env python
"""Per-phase timeline of K1 (experiment builds with -DFQD_K1_TIMELINE only): runs a few chunks of synthetic FASTQ
through the --fast path and prints, over the sampled tiles of the last launch, the median / p90 clock64 deltas
between the stamps in parse_pack.cuh (SM cycles; 1965 MHz when not throttled)."""
import ctypes as C
import importlib
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
fqd = importlib.import_module("fastq-dupaway_b200")
lib = fqd.load_library()
n_reads, chunk = 18_000_000, 6_000_000
REC = 322
raw = fqd.DeviceBuffer(n_reads * REC + 65536, 0)
for c in range(n_reads // chunk):
    assert lib.fqd_synth_fastq(0, raw.ptr + c * chunk * REC, c * chunk, chunk, 150, 1, 1, 300, 1, 0) == 0
eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, False, False, 2, 150, n_reads + 1024, chunk * REC + 65536, chunk + 1024, 0)
for _ in range(3):
    eng.reset()
    for c in range(n_reads // chunk):
        eng.push_device_async(raw.ptr + c * chunk * REC, chunk * REC)
eng.sync()
SLOTS, CAP = 12, 4096
buf = (C.c_longlong * (SLOTS * CAP))()
lib.fqd_debug_k1_timeline.argtypes = [C.c_void_p, C.c_size_t]
assert lib.fqd_debug_k1_timeline(buf, SLOTS * CAP) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(CAP, SLOTS)
n_tiles = (chunk * REC + 16383) // 16384
t = t[: min(CAP, n_tiles // 61)]
t = t[(t[:, 0] > 0) & (t[:, 8] > 0)]
names = ["start", "tma+sync", "mask sync", "scan+publish", "pos barrier (packers)", "owners (BAR_WORK)", "pack done", "P known (BAR_P)", "commit done",
         "lookback start (w0)", "lookback end (w0)"]
def stat(a):
    return f"median {np.median(a):8.0f}  p10 {np.percentile(a, 10):8.0f}  p90 {np.percentile(a, 90):8.0f}"
print(f"{len(t)} sampled tiles; SM cycles since CTA start")
for k in range(1, 11):
    print(f"  {names[k]:28s} {stat(t[:, k] - t[:, 0])}")
print("phase lengths")
for a, b in [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (9, 10), (3, 9), (10, 7)]:
    print(f"  {names[a]:28s} -> {names[b]:28s} {stat(t[:, b] - t[:, a])}")
