"""Seeded parity cases: small versions of every BASELINE.json config plus edge cases.

Shared by tests/golden/make_golden.py (which runs the UNMODIFIED reference on the
SAM view of each case, in the build container) and by the tests (which rebuild
the interval view from the same seeds and never touch /root/reference)."""
from __future__ import annotations

from dataclasses import dataclass, field

L3 = [300000, 200000, 100000]
BED3 = [(0, 50000, 61000), (0, 0, 1500), (0, 60500, 70000), (0, 70000, 70500), (0, 150000, 150001), (0, 299000, 400000),
        (1, 19900, 20100), (1, 39000, 41000), (1, 40000, 40500), (1, 250000, 260000), (2, 99999, 100000), (2, 8191, 8193)]


@dataclass
class Sample:
    nfrag: int
    seed: int
    enrich: float = 0.3
    spacing: int = 20000
    sigma: float = 100.0
    multimap: float = 0.0
    mmax: int = 12
    drop_chroms: tuple = ()      # chromosomes (0-based) removed from this file's header and reads
    empty_chroms: tuple = ()     # chromosomes whose reads are removed, header kept (Chrom.diff stays NULL)


@dataclass
class Case:
    name: str
    chrom_len: list
    reps: list                   # list of (expt Sample, ctrl Sample | None)
    p: float | None = None
    q: float | None = None
    min_auc: float = 200.0
    min_len: int = 0
    max_gap: int = 100
    atac: bool = False
    atac_len: int = 100
    as_diff: float = 0.0
    extra: list = field(default_factory=list)
    bed: list = field(default_factory=list)   # -E regions: (chromosome index, start, end), as a BED file would list them

    def ref_args(self):
        a = []
        if self.q is not None:
            a += ["-q", repr(self.q)]
        elif self.p is not None:
            a += ["-p", repr(self.p)]
        if self.min_auc != 200.0:
            a += ["-a", repr(self.min_auc)]
        if self.min_len:
            a += ["-l", str(self.min_len)]
        if self.max_gap != 100:
            a += ["-g", str(self.max_gap)]
        if self.atac:
            a += ["-j", "-d", str(self.atac_len)]
        if self.as_diff:
            a += ["-s", repr(self.as_diff)]
        return a + self.extra


CASES = [
    # BASELINE config 1, exactly
    Case("c1_smoke", [1000000], [(Sample(100000, 1001, enrich=0.2, spacing=50000, sigma=150.0), None)], p=0.01),
    # config 2 in miniature: treatment + control, default thresholds but -q so BH is exercised too
    Case("c2_ctrl_q", L3, [(Sample(60000, 21), Sample(60000, 22, enrich=0.0))], q=0.05),
    Case("c2_ctrl_p", L3, [(Sample(60000, 21), Sample(60000, 22, enrich=0.0))], p=0.01),
    # config 3: ATAC mode (clamps at both chromosome ends occur)
    Case("c3_atac_q", L3, [(Sample(60000, 31, enrich=0.4, sigma=40.0), None)], q=0.05, atac=True),
    # config 4: three replicates, one control, Fisher combine
    Case("c4_fisher_q", L3, [(Sample(60000, 21), Sample(60000, 22, enrich=0.0)),
                             (Sample(60000, 23), None), (Sample(60000, 24), None)], q=0.05),
    # config 5: multimapping fractional weights
    Case("c5_multimap_q", [300000, 200000], [(Sample(60000, 25, multimap=0.3), None)], q=0.05, as_diff=20.0),
    Case("c5_multimap_ctrl_p", [300000, 200000],
         [(Sample(40000, 26, multimap=0.3), Sample(40000, 27, enrich=0.0, multimap=0.3))], p=0.01, as_diff=20.0),
    # peak-shape parameters
    Case("p_gap_len_auc", L3, [(Sample(60000, 21), None)], p=0.05, min_auc=50.0, min_len=120, max_gap=20),
    Case("p_gap0", L3, [(Sample(60000, 21), None)], p=0.001, min_auc=1.0, max_gap=0),
    # sparse coverage: many empty stretches, one chromosome with no reads at all in the control
    Case("sparse_ctrl", [500000, 40000, 30000], [(Sample(3000, 41, enrich=0.5, spacing=25000, sigma=60.0),
                                                  Sample(300, 42, enrich=0.0, drop_chroms=()))], p=0.01, min_auc=20.0),
    # a replicate whose file lacks a chromosome (Chrom.save false -> NULL p array, Genrich.c:1744)
    Case("fisher_missing_chrom", L3, [(Sample(60000, 21), None), (Sample(60000, 23, drop_chroms=(2,)), None)], q=0.05),
    # all q-values are 1 (no enrichment) -> zero peaks, warning path
    Case("null_q", [200000], [(Sample(20000, 51, enrich=0.0), None)], q=0.05),
    # -E excluded regions: overlapping / touching / unsorted records, a region at 0, one running past the
    # chromosome end, one starting beyond it (ignored), regions cutting through peaks (spacing 20000)
    Case("bed_ctrl_q", L3, [(Sample(60000, 21), Sample(60000, 22, enrich=0.0))], q=0.05, bed=BED3),
    Case("bed_noctrl_p", L3, [(Sample(60000, 21), None)], p=0.01, bed=BED3),
    Case("bed_fisher_q", L3, [(Sample(60000, 21), Sample(60000, 22, enrich=0.0)),
                              (Sample(60000, 23, drop_chroms=(2,)), None)], q=0.05, bed=BED3),
    # a chromosome that never receives a read, with regions on it (saveLambda 1847-1877 / saveConst 2178)
    Case("bed_empty_chrom", [300000, 50000], [(Sample(30000, 61, empty_chroms=(1,)), Sample(30000, 62, enrich=0.0, empty_chroms=(1,)))],
         p=0.01, bed=[(1, 0, 1000), (1, 20000, 60000), (0, 100, 200)]),
    # ... and one that is read-less in both experimental samples but not in the first control: the reference's diff
    # array (shared by all samples) exists from then on, so the SECOND experimental pileup is cut at the region
    # boundaries (the normal loop of savePileupExpt) while the first is one interval (saveConst 2178)
    Case("bed_later_empty", [300000, 50000], [(Sample(30000, 71, empty_chroms=(1,)), Sample(30000, 72, enrich=0.0)),
                                              (Sample(30000, 73, empty_chroms=(1,)), None)],
         p=0.01, bed=[(1, 0, 1000), (1, 20000, 60000), (0, 100, 200)]),
]

BY_NAME = {c.name: c for c in CASES}
