#!/usr/bin/env python
"""Generate tests/golden/host_* by running the UNMODIFIED reference (oracle/_ref/Genrich) on the
mutated SAM files of tests/hostcases.py (unpaired / discordant alignments, PCR duplicates, quality
strings, low MAPQ) with the host-side options of each case.  Run in the build container only.

  host_<case>.narrowPeak   the reference's -o file, verbatim (absent for -X)
  host_<case>.json         sha256 + line count of its -f log, its -b interval file and its -R duplicates log (without
                           the '# ... file' lines, which hold temporary paths), and the complete -v
                           text with the temporary directory replaced by '@'
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hostcases import HOST_CASES, write_host_sams, host_cmd  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "Genrich")


def sha(path, skip_hash=False):
    h = hashlib.sha256()
    n = 0
    with open(path, "rb") as f:
        for line in f:
            if skip_hash and line.startswith(b"#"):
                continue
            h.update(line)
            n += 1
    return h.hexdigest(), n


def main():
    only = set(sys.argv[1:])
    for hc in HOST_CASES:
        if only and hc.name not in only:
            continue
        with tempfile.TemporaryDirectory() as td:
            tfiles, cfiles = write_host_sams(hc, td)
            bedf = os.path.join(td, "o.bed")
            cmd, out, logf, dupf = host_cmd(REF, hc, td, tfiles, cfiles, bed=bedf)
            r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
            if r.returncode != 0:
                raise SystemExit("reference failed on %s:\n%s" % (hc.name, r.stderr))
            meta = {"args": [x for x in cmd[1:][cmd[1:].index("-v"):] if x not in ("-b", bedf)], "stderr": r.stderr.replace(td, "@")}
            meta["log_sha256"], meta["log_lines"] = sha(logf)
            meta["bed_sha256"], meta["bed_lines"] = sha(bedf)
            if hc.dups_log:
                meta["dups_sha256"], meta["dups_lines"] = sha(dupf, True)
            if os.path.exists(out):
                with open(out) as f, open(os.path.join(HERE, hc.name + ".narrowPeak"), "w") as g:
                    g.write(f.read())
                meta["peaks"] = sum(1 for _ in open(out))
            with open(os.path.join(HERE, hc.name + ".json"), "w") as f:
                json.dump(meta, f, indent=1, sort_keys=True)
            print(hc.name, meta.get("peaks"), "peaks", meta["log_lines"], "log lines", meta.get("dups_lines"), "dup lines")


if __name__ == "__main__":
    main()
