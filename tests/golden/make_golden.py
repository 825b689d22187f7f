#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (oracle/_ref/Genrich,
built by `make -C oracle ref` from /root/reference) on the SAM view of every case
in tests/cases.py.  Run in the build container only; the outputs are committed:

  <case>.narrowPeak   the reference's -o file, verbatim
  <case>.json         sha256 + line count of its -f log (and -k pileup file),
                      the -v scalars (lambda, scale factor, genome length, peaks)

The -f/-k text is pinned by hash, not stored: the oracle must reproduce it byte
for byte (tests/test_oracle_pin.py), which pins every interval boundary, pileup
value, -log10(p) and -log10(q) to the 6 decimals the reference prints.
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import CASES  # noqa: E402
import util  # noqa: E402
from genrich_b200.synth import Workload  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "Genrich")


def sha(path):
    h = hashlib.sha256()
    n = 0
    with open(path, "rb") as f:
        for line in f:
            h.update(line)
            n += 1
    return h.hexdigest(), n


def main():
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case.name not in only:
            continue
        with tempfile.TemporaryDirectory() as td:
            tfiles, cfiles = util.write_case_sams(case, td)
            out = os.path.join(HERE, case.name + ".narrowPeak")
            logf = os.path.join(td, "log.f")
            pile = os.path.join(td, "pile.k")
            cmd = [REF, "-t", ",".join(tfiles), "-o", out, "-f", logf, "-k", pile, "-v"] + case.ref_args()
            if any(c != "null" for c in cfiles):
                cmd += ["-c", ",".join(cfiles)]
            if case.bed:
                bedf = os.path.join(td, "x.bed")
                util.write_case_bed(case, bedf)
                cmd += ["-E", bedf]
            r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
            if r.returncode != 0:
                raise SystemExit("reference failed on %s:\n%s" % (case.name, r.stderr))
            err = r.stderr
            meta = {
                "args": case.ref_args(),
                "lambda": [float(x) for x in re.findall(r"Background pileup value: ([0-9.]+)", err)],
                "factor": [float(x) for x in re.findall(r"Scaling factor for control pileup: ([0-9.]+)", err)],
                "genome_len": int(re.search(r"Genome length: (\d+)bp", err).group(1)),
                "peaks": int(re.search(r"Peaks identified: (\d+)", err).group(1)),
                "peak_bp": int(re.search(r"Peaks identified: \d+ \((\d+)bp\)", err).group(1)),
                "all_q_one": "All q-values are 1" in err,
                "clamp_warnings": len(re.findall(r"prevented from extending", err)),
            }
            meta["log_sha256"], meta["log_lines"] = sha(logf)
            # -k: strip the '# experimental file: ...' lines (they hold temp paths)
            kept = os.path.join(td, "pile.nohdr")
            with open(pile) as f, open(kept, "w") as g:
                for line in f:
                    if not line.startswith("#"):
                        g.write(line)
            meta["pile_sha256"], meta["pile_lines"] = sha(kept)
            with open(os.path.join(HERE, case.name + ".json"), "w") as f:
                json.dump(meta, f, indent=1, sort_keys=True)
            print(case.name, meta["peaks"], "peaks", meta["log_lines"], "log lines")


if __name__ == "__main__":
    main()
