"""Seeded RANDOM cases (test infrastructure): small chromosome tables with awkward lengths, 1-3 replicates with or
without control, -p / -q, gap / length / AUC thresholds, multimapped weights, ATAC intervals of odd lengths, -E regions
(out of range, overlapping, on excluded chromosomes), read-less and header-less chromosomes; and random host options.

Used three ways: the CUDA library compiled for the CPU against the oracle (tests/test_emu_library.py), the host
program against the unmodified reference binary on the same SAM files (tests/test_cli_host.py), and -- offline, over
hundreds of seeds -- the oracle against the reference binary.  What they found so far: the reference's delta array is
shared by all samples of a run (a read-less chromosome is only one interval until some sample had reads there), the
order and the conditions of saveXBed's -v warnings, the point at which an empty experimental sample ends the run."""
from __future__ import annotations

import random

import numpy as np

from cases import Case, Sample


def random_case(seed: int, holes_with_multimap: bool = True) -> Case:
    """holes_with_multimap=False: no multimapped templates in cases with read-less / header-less chromosomes (the
    interval view of cases.py drops such placements AFTER the weights are fixed, the SAM view before: only the SAM
    view is what the reference sees)"""
    r = np.random.RandomState(seed)
    nchrom = int(r.randint(1, 6))
    L = [int(x) for x in r.choice([300, 8191, 8192, 8193, 16384, 20000, 70000, 150000], nchrom)]
    if max(L) < 20000:
        L[int(r.randint(nchrom))] = 90000
    nrep = int(r.choice([1, 2, 3]))
    reps = []
    for k in range(nrep):
        drop = ()
        if nchrom > 1 and k > 0 and r.uniform() < 0.3:
            drop = (int(r.randint(nchrom)),)
        emp = (int(r.randint(nchrom)),) if nchrom > 1 and r.uniform() < 0.4 else ()
        e = Sample(int(r.randint(500, 9000)), 100 * seed + k, enrich=float(r.choice([0.0, 0.2, 0.6])),
                   spacing=int(r.choice([2000, 5000, 20000])), sigma=float(r.choice([5.0, 30.0, 100.0])),
                   multimap=float(r.choice([0.0, 0.3, 0.8])), drop_chroms=drop, empty_chroms=emp)
        cemp = (int(r.randint(nchrom)),) if nchrom > 1 and r.uniform() < 0.3 else ()
        c = Sample(int(r.randint(500, 9000)), 100 * seed + 50 + k, enrich=float(r.choice([0.0, 0.1])),
                   multimap=float(r.choice([0.0, 0.3])), empty_chroms=cemp, drop_chroms=drop) if r.uniform() < 0.6 else None
        reps.append((e, c))
    use_q = r.uniform() < 0.5
    bed = []
    if r.uniform() < 0.5:
        for _ in range(int(r.randint(1, 8))):
            c = int(r.randint(nchrom))
            s = int(r.randint(0, max(L[c], 2)))
            bed.append((c, s, s + int(r.choice([1, 2, 50, 5000, 200000]))))
    holes = any(e.drop_chroms or e.empty_chroms or (c is not None and (c.drop_chroms or c.empty_chroms)) for e, c in reps)
    if holes and not holes_with_multimap:
        for e, c in reps:
            e.multimap = 0.0
            if c is not None:
                c.multimap = 0.0
    mm = any(e.multimap or (c is not None and c.multimap) for e, c in reps)
    return Case("fuzz%d" % seed, L, reps, p=None if use_q else float(r.choice([0.01, 0.05, 0.5])),
                q=float(r.choice([0.05, 0.5, 0.9])) if use_q else None, min_auc=float(r.choice([0.0, 20.0, 200.0])),
                min_len=int(r.choice([0, 0, 150])), max_gap=int(r.choice([0, 1, 100, 1000])),
                atac=bool(r.uniform() < 0.25), atac_len=int(r.choice([30, 100, 501])), as_diff=20.0 if mm else 0.0, bed=bed)


def random_host_options(seed: int, case: Case) -> list:
    r = random.Random(seed)
    ex = []
    k = r.choice(["", "-y", "-w", "-x", "-y", ""])
    if k == "-w":
        ex += ["-w", str(r.choice([50, 150, 300]))]
    elif k and not (k == "-x" and case.atac):
        ex += [k]
    if r.random() < 0.4:
        ex += ["-r"]
    if r.random() < 0.3:
        ex += ["-m", str(r.choice([5, 15, 30]))]
    if r.random() < 0.2 and len(case.chrom_len) > 1:
        ex += ["-e", "chr%d" % (r.randrange(len(case.chrom_len)) + 1)]
    if r.random() < 0.2:
        ex += ["-D"]
    if r.random() < 0.15:
        ex += ["-X"]
    if r.random() < 0.2:
        ex += ["-S"]
    return ex
