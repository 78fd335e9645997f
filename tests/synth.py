"""Seeded synthetic FASTA/FASTQ inputs for the parity tests (numpy; small sizes)."""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_reads(n, seed=1, read_len=50, var_len=False, dup_frac=0.3, n_frac=0.0, prefix_frac=0.0,
               sub_frac=0.0, alphabet=b"ACGT", min_len=0):
    """Return a list of n sequences (bytes).  dup_frac of the reads copy an earlier read (chains allowed);
    prefix_frac of those are truncated (loose-mode material), sub_frac get 1-2 tail substitutions."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    seqs = []
    for i in range(n):
        if i > 0 and rng.random() < dup_frac:
            s = bytearray(seqs[int(rng.integers(0, i))])
            r = rng.random()
            if r < prefix_frac and len(s) > min_len + 1:
                s = s[: len(s) - int(rng.integers(1, min(10, len(s) - min_len) + 1))]
            elif r < prefix_frac + sub_frac and len(s) > 0:
                for _ in range(int(rng.integers(1, 3))):
                    p = len(s) - 1 - int(rng.integers(0, min(10, len(s))))
                    s[p] = int(alpha[int(rng.integers(0, len(alpha)))])
            seqs.append(bytes(s))
            continue
        ln = int(rng.integers(min_len, read_len + 1)) if var_len else read_len
        s = alpha[rng.integers(0, len(alpha), size=ln)].copy()
        if n_frac > 0 and ln > 0 and rng.random() < n_frac:
            s[int(rng.integers(0, ln))] = ord("N")
        seqs.append(s.tobytes())
    return seqs


def to_fastq(seqs, seed=7, mate=1, id_fmt="@SYN.{i:010d} {mate}", ids=None, plus=b"+"):
    rng = np.random.default_rng(seed)
    q = np.frombuffer(b"FGHI", dtype=np.uint8)
    out = []
    for i, s in enumerate(seqs):
        ident = ids[i] if ids is not None else id_fmt.format(i=i, mate=mate).encode()
        qual = q[rng.integers(0, 4, size=len(s))].tobytes()
        out.append(ident + b"\n" + s + b"\n" + plus + b"\n" + qual + b"\n")
    return b"".join(out)


def to_fasta(seqs, mate=1, id_fmt=">SYN.{i:010d} {mate}", ids=None):
    out = []
    for i, s in enumerate(seqs):
        ident = ids[i] if ids is not None else id_fmt.format(i=i, mate=mate).encode()
        out.append(ident + b"\n" + s + b"\n")
    return b"".join(out)


def make_pair(n, seed=1, **kw):
    """Paired reads: duplicates copy BOTH mates of an earlier pair (with some pairs differing in R2 only)."""
    rng = np.random.default_rng(seed + 1000)
    r1 = make_reads(n, seed=seed, dup_frac=0.0, **{k: v for k, v in kw.items() if k != "dup_frac"})
    r2 = make_reads(n, seed=seed + 1, dup_frac=0.0, **{k: v for k, v in kw.items() if k != "dup_frac"})
    dup_frac = kw.get("dup_frac", 0.3)
    for i in range(1, n):
        u = rng.random()
        if u < dup_frac:
            j = int(rng.integers(0, i))
            r1[i], r2[i] = r1[j], r2[j]
        elif u < dup_frac + 0.05:
            j = int(rng.integers(0, i))
            r1[i] = r1[j]          # same R1, different R2 -> NOT a duplicate pair
    return r1, r2
