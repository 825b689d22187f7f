"""The int16 saturation case (saveInterval, Genrich.c:2558-2573): a seeded sample whose hot spots take the
reference's per-base delta counters to INT16_MAX / INT16_MIN, so that it SKIPS intervals in arrival order.

Shared by tests/golden/make_golden_sat.py (runs the unmodified reference on the SAM view) and by the
tests (oracle and CUDA path on the interval view).  One chromosome, templates in arrival order:

  A  40 000 fragments that START on one base (ends spread out)           -> overflow after 32 767
  B  40 000 fragments that END on one base (starts spread out)           -> underflow after 32 768
  C  70 000 placements of weight 1/2 starting on one base, interleaved with fragments that END on that
     base (they lower the counter again, so WHICH starts are dropped depends on the order)
  D  fragments from hot base A to hot base B' (an interval dropped for its start never reaches its end)
  +  an ordinary background so that peaks are called around the hot spots
"""
import numpy as np

from genrich_b200.synth import Workload

CHROM_LEN = [400000]
ARGS = ["-p", "0.01", "-s", "20"]
A, B, C, B2 = 50000, 120000, 150000, 50200


def templates():
    """list of templates in file order; a template = list of (start, end) placements (k placements -> weight
    1/k each, k in {1, 2})"""
    rng = np.random.RandomState(7)
    t = []
    for i in range(40000):
        t.append([(A, A + 100 + i % 300)])
    for i in range(40000):
        t.append([(B - 100 - i % 250, B)])
    for i in range(35000):
        t.append([(C, C + 120 + i % 200), (C, C + 130 + i % 170)])       # two placements, both starting on C
    for i in range(3000):
        t.append([(C - 150 - i % 100, C)])                              # ends on C
    for i in range(34000):
        t.append([(A, B2)])                                             # A -> B2: B2 collects ends
    for i in range(2000):
        t.append([(B2 - 120 - i % 50, B2)])
    bg = Workload(CHROM_LEN, 30000, 77, enrich=0.3, spacing=20000, sigma=100.0).fragments()
    for r in bg:
        t.append([(int(r[1]), int(r[2]))])
    order = rng.permutation(len(t))
    return [t[i] for i in order]


def records():
    """interval records in arrival order: (chrom, start, end, count)"""
    out = []
    for pl in templates():
        for s, e in pl:
            out.append((0, s, e, len(pl)))
    return np.array(out, dtype=np.int32)


def write_sam(path, read_len=50):
    """queryname-sorted SAM as gen_synth writes it: proper pairs 99 / 147, secondary placements + 256, AS:i:0"""
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:chr1\tLN:%d\n" % CHROM_LEN[0])
        for n, pl in enumerate(templates()):
            for i, (s, e) in enumerate(pl):
                sec = 256 if i else 0
                r2 = max(e - read_len, 0)
                f.write("f%d\t%d\tchr1\t%d\t42\t%dM\t=\t%d\t%d\t*\t*\tAS:i:0\n" % (n, 99 + sec, s + 1, read_len, r2 + 1, e - s))
                f.write("f%d\t%d\tchr1\t%d\t42\t%dM\t=\t%d\t%d\t*\t*\tAS:i:0\n" % (n, 147 + sec, r2 + 1, read_len, s + 1, -(e - s)))
