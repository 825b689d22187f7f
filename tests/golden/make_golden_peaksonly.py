#!/usr/bin/env python
"""Generate tests/golden/ponly_* by running the UNMODIFIED reference (oracle/_ref/Genrich) in -P mode
(peaks from a -f log, callPeaksLog Genrich.c:1277) on logs it wrote itself for seeded cases.  Run in
the build container only; the outputs are committed:

  ponly_<name>.narrowPeak   the reference's -o file, verbatim
  ponly_<name>.json         its -v scalars (genome length, peaks, bp) and the arguments

tests/test_cli_host.py::test_peaks_only replays the same commands with the host program (whose own -f
log is the reference's byte for byte, tests/golden/<case>.json log_sha256)."""
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

from cases import BY_NAME  # noqa: E402
import util  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "Genrich")

# name -> (case whose -f log is used, -P arguments, new -E regions from this case or None)
PONLY = {
    "ponly_c2_same": ("c2_ctrl_q", ["-q", "0.05"], None),
    "ponly_c2_p_strict": ("c2_ctrl_q", ["-p", "0.001", "-a", "50", "-l", "50", "-g", "30"], None),
    "ponly_c2_q_skipchr": ("c2_ctrl_q", ["-q", "0.2", "-e", "chr2", "-a", "20"], None),
    "ponly_c2_newbed": ("c2_ctrl_q", ["-p", "0.01", "-a", "100"], "bed_ctrl_q"),
    "ponly_c2_newbed_q": ("c2_ctrl_q", ["-q", "0.1", "-g", "0", "-a", "1"], "bed_ctrl_q"),
    "ponly_fisher": ("c4_fisher_q", ["-q", "0.05"], None),
    "ponly_fisher_p": ("c4_fisher_q", ["-p", "0.0001", "-L", "1000000"], None),
    "ponly_bedlog": ("bed_ctrl_q", ["-p", "0.01"], None),
    "ponly_bedlog_newbed": ("bed_fisher_q", ["-q", "0.05", "-a", "50"], "bed_ctrl_q"),
    "ponly_smoke_p": ("c1_smoke", ["-p", "0.05", "-l", "200"], None),
    "ponly_atac": ("c3_atac_q", ["-q", "0.01", "-g", "250"], None),
}


def ref_log(case, td):
    tfiles, cfiles = util.write_case_sams(case, td)
    out, logf = os.path.join(td, "o.np"), os.path.join(td, "log.f")
    cmd = [REF, "-t", ",".join(tfiles), "-o", out, "-f", logf] + case.ref_args()
    if any(c != "null" for c in cfiles):
        cmd += ["-c", ",".join(cfiles)]
    if case.bed:
        bedf = os.path.join(td, "x.bed")
        util.write_case_bed(case, bedf)
        cmd += ["-E", bedf]
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)
    return logf


def main():
    for name, (cname, args, bedcase) in PONLY.items():
        with tempfile.TemporaryDirectory() as td:
            logf = ref_log(BY_NAME[cname], td)
            out = os.path.join(HERE, name + ".narrowPeak")
            cmd = [REF, "-P", "-f", logf, "-o", out, "-v"] + args
            if bedcase:
                bedf = os.path.join(td, "new.bed")
                util.write_case_bed(BY_NAME[bedcase], bedf)
                cmd += ["-E", bedf]
            r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
            if r.returncode != 0:
                raise SystemExit("reference -P failed on %s:\n%s" % (name, r.stderr))
            err = r.stderr
            meta = {"case": cname, "args": args, "bed_case": bedcase,
                    "genome_len": int(re.search(r"Genome length: (\d+)bp", err).group(1)),
                    "peaks": int(re.search(r"Peaks identified: (\d+)", err).group(1)),
                    "peak_bp": int(re.search(r"Peaks identified: \d+ \((\d+)bp\)", err).group(1)),
                    "warn_bed": "Skipping given BED regions" in err,
                    "warn_chr": len(re.findall(r"Skipping chromosome", err))}
            with open(os.path.join(HERE, name + ".json"), "w") as f:
                json.dump(meta, f, indent=1, sort_keys=True)
            print(name, meta["peaks"], "peaks", meta["genome_len"], "bp")


if __name__ == "__main__":
    main()
