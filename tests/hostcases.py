"""Host-side option cases: SAM files that exercise what the seeded fragment sets of cases.py do
not contain -- unpaired and discordant alignments, PCR duplicates, quality strings, low MAPQ --
so that -y / -w / -x / -r / -R / -m / -e / -X of the host program can be compared with the
unmodified reference (tests/golden/host_*.{narrowPeak,json}, written by make_golden_host.py)."""
from __future__ import annotations

import os
import random
from dataclasses import dataclass, field

import util
from cases import Case, Sample, L3


def mutate_sam(src: str, dst: str, seed: int) -> None:
    """Rewrite a proper-pairs SAM file (queryname-grouped) deterministically:
    every read gets quality strings; of the two-line reads, by index i:
      i % 10 == 0  both mates lose the proper-pair bit            -> a discordant set
      i % 10 == 1  mate 2 is dropped, mate 1 is not a proper pair -> a singleton (R1)
      i % 10 == 2  mate 1 is dropped                              -> a singleton (R2)
      i % 10 == 3  MAPQ 10 on both mates
      i %  7 == 4  the read is repeated under a new name right after itself (a PCR duplicate
                   with its own qualities; the discordant and singleton sets are repeated too)
    Reads with more than two lines (multimappers) only get qualities."""
    rng = random.Random(seed)
    out = []
    group = []

    def qual():
        return "".join(chr(33 + rng.randrange(2, 41)) for _ in range(50))

    def flush(i):
        if not group:
            return
        lines = [l.split("\t") for l in group]
        q1, q2 = qual(), qual()
        for f in lines:
            f[10] = q1 if int(f[1]) & 0x40 else q2
        if len(lines) == 2:
            k = i % 10
            if k == 0:
                for f in lines:
                    f[1] = str(int(f[1]) & ~0x2)
            elif k == 1:
                lines = [f for f in lines if int(f[1]) & 0x40]
                lines[0][1] = str(int(lines[0][1]) & ~0x2)
            elif k == 2:
                lines = [f for f in lines if int(f[1]) & 0x80]
                lines[0][1] = str(int(lines[0][1]) & ~0x2)
            elif k == 3:
                for f in lines:
                    f[4] = "10"
        out.extend("\t".join(f) for f in lines)
        if len(group) == 2 and i % 7 == 4:
            q1, q2 = qual(), qual()
            for f in lines:
                g = list(f)
                g[0] = "dup_" + f[0]
                g[10] = q1 if int(f[1]) & 0x40 else q2
                out.append("\t".join(g))

    i = 0
    name = None
    with open(src) as f:
        for line in f:
            if line.startswith("@"):
                out.append(line.rstrip("\n"))
                continue
            line = line.rstrip("\n")
            q = line.split("\t", 1)[0]
            if q != name:
                flush(i)
                if name is not None:
                    i += 1
                group = []
                name = q
            group.append(line)
    flush(i)
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


@dataclass
class HostCase:
    name: str
    case: Case                    # fragment sets and peak-calling options
    args: list = field(default_factory=list)     # extra command-line options (host side)
    dups_log: bool = False        # also write / compare the -R file


_T = Sample(30000, 71, enrich=0.3)
_C = Sample(30000, 72, enrich=0.0)
_M = Sample(20000, 73, enrich=0.3, multimap=0.3)

HOST_CASES = [
    HostCase("host_y", Case("host_y", L3, [(_T, None)], p=0.01), ["-y"]),
    HostCase("host_w", Case("host_w", L3, [(_T, _C)], p=0.01), ["-w", "180"]),
    HostCase("host_x", Case("host_x", L3, [(_T, None)], p=0.01), ["-x"]),
    HostCase("host_m_e", Case("host_m_e", L3, [(_T, None)], p=0.01), ["-y", "-m", "20", "-e", "chr2"]),
    HostCase("host_X", Case("host_X", L3, [(_T, None)], q=0.05), ["-X"]),
    HostCase("host_r", Case("host_r", L3, [(_T, _C)], p=0.01), ["-r"], dups_log=True),
    HostCase("host_r_y", Case("host_r_y", L3, [(_T, None)], p=0.01), ["-r", "-y"], dups_log=True),
    HostCase("host_r_x", Case("host_r_x", L3, [(_T, _C)], p=0.01), ["-r", "-x"], dups_log=True),
    HostCase("host_r_w", Case("host_r_w", L3, [(_T, None)], p=0.01), ["-r", "-w", "150"]),
    HostCase("host_r_atac", Case("host_r_atac", L3, [(_T, None)], p=0.01, atac=True), ["-r", "-y"], dups_log=True),
    HostCase("host_r_multimap", Case("host_r_multimap", [300000, 200000], [(_M, None)], p=0.01, as_diff=20.0),
             ["-r", "-y"], dups_log=True),
]
HOST_BY_NAME = {h.name: h for h in HOST_CASES}


def write_host_sams(h: HostCase, td: str):
    """SAM files of the case, mutated; returns (treatment files, control files or 'null')."""
    tfiles, cfiles = util.write_case_sams(h.case, td)
    outs_t, outs_c = [], []
    for k, p in enumerate(tfiles):
        q = os.path.join(td, "mt%d.sam" % k)
        mutate_sam(p, q, 500 + k)
        outs_t.append(q)
    for k, p in enumerate(cfiles):
        if p == "null":
            outs_c.append(p)
            continue
        q = os.path.join(td, "mc%d.sam" % k)
        mutate_sam(p, q, 600 + k)
        outs_c.append(q)
    return outs_t, outs_c


def host_cmd(binary: str, h: HostCase, td: str, tfiles, cfiles, bed=None):
    out, logf, dupf = (os.path.join(td, x) for x in ("o.np", "o.f", "o.R"))
    cmd = [binary, "-t", ",".join(tfiles), "-f", logf, "-v"] + h.case.ref_args() + list(h.args)
    if bed:
        cmd += ["-b", bed]
    if "-X" not in h.args:
        cmd += ["-o", out]
    if any(c != "null" for c in cfiles):
        cmd += ["-c", ",".join(cfiles)]
    if h.dups_log:
        cmd += ["-R", dupf]
    return cmd, out, logf, dupf
