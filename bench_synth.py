"""CPU (numpy) twin of the device generator in fastq-dupaway_b200/csrc/synth.cuh, byte for byte (variant 0).
Used by bench.py's CPU legs (cpu_baseline, --impl reference) and by tests to cross-check the device generator."""
from __future__ import annotations

import numpy as np

U = np.uint64
C1, C2, C3 = U(0x9E3779B97F4A7C15), U(0xBF58476D1CE4E5B9), U(0x94D049BB133111EB)
K = U(0xD1342543DE82EF95)


def _splitmix(x):
    x = x + C1
    x = (x ^ (x >> U(30))) * C2
    x = (x ^ (x >> U(27))) * C3
    return x ^ (x >> U(31))


def _rng(seed, i, stream):
    s = _splitmix(np.array([seed], dtype=U) ^ (np.array([stream], dtype=U) * K))
    return _splitmix(s + i * C1)


def synth_roots(first, count, seed, dup_permille):
    with np.errstate(over="ignore"):
        cur = np.arange(first, first + count, dtype=U)
        while True:
            r = _rng(seed, cur, 1)
            is_dup = ((r % U(1000)) < U(dup_permille)) & (cur > U(0))
            if not is_dup.any():
                return cur
            src = (r >> U(20)) % np.maximum(cur, U(1))
            cur = np.where(is_dup, src, cur)


def synth_fastq_cpu(first, count, read_len, mate, seed, dup_permille, n_permille) -> bytes:
    L = read_len
    rec = 22 + 2 * L
    out = np.empty((count, rec), dtype=np.uint8)
    with np.errstate(over="ignore"):
        idx = np.arange(first, first + count, dtype=U)
        root = synth_roots(first, count, seed, dup_permille)
        out[:, 0:5] = np.frombuffer(b"@SYN.", dtype=np.uint8)
        v = idx.copy()
        for d in range(10):
            out[:, 14 - d] = (v % U(10)).astype(np.uint8) + ord("0")
            v //= U(10)
        out[:, 15] = ord(" ")
        out[:, 16] = ord("0") + mate
        out[:, 17] = ord("\n")
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        fghi = np.frombuffer(b"FGHI", dtype=np.uint8)
        sh = (U(2) * np.arange(32, dtype=U))[None, :]
        nblk = (L + 31) // 32
        for b in range(nblk):
            w = min(32, L - 32 * b)
            rb = _rng(seed, root * U(64) + U(b), 16 + mate)
            out[:, 18 + 32 * b: 18 + 32 * b + w] = acgt[((rb[:, None] >> sh[:, :w]) & U(3)).astype(np.intp)]
            rq = _rng(seed, idx * U(16) + U(b), 32 + mate)
            q0 = 18 + L + 3 + 32 * b
            out[:, q0: q0 + w] = fghi[((rq[:, None] >> sh[:, :w]) & U(3)).astype(np.intp)]
        rn = _rng(seed, root, 8 + mate)
        has_n = (rn % U(1000)) < U(n_permille)
        npos = ((rn >> U(20)) % U(L)).astype(np.intp)
        rows = np.flatnonzero(has_n)
        out[rows, 18 + npos[rows]] = ord("N")
        out[:, 18 + L] = ord("\n")
        out[:, 19 + L] = ord("+")
        out[:, 20 + L] = ord("\n")
        out[:, rec - 1] = ord("\n")
    return out.tobytes()
