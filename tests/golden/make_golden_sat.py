#!/usr/bin/env python
"""Golden files of the int16 saturation case (tests/satcase.py): the UNMODIFIED reference on its SAM view.

  sat_hot.narrowPeak   the reference's -o file, verbatim
  sat_hot.json         sha256 + line count of its -f / -k text, lambda, and how many intervals it reported
                       "skipped due to overflow" / "... underflow" (saveInterval, Genrich.c:2558-2573)
Run in the build container only (needs oracle/_ref/Genrich)."""
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
import satcase  # noqa: E402

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
    with tempfile.TemporaryDirectory() as td:
        sam = os.path.join(td, "t.sam")
        satcase.write_sam(sam)
        out = os.path.join(HERE, "sat_hot.narrowPeak")
        logf, pile = os.path.join(td, "log.f"), os.path.join(td, "pile.k")
        r = subprocess.run([REF, "-t", sam, "-o", out, "-f", logf, "-k", pile, "-v"] + satcase.ARGS, stderr=subprocess.PIPE, text=True)
        if r.returncode:
            raise SystemExit(r.stderr[-2000:])
        err = r.stderr
        over = re.findall(r"Warning! Read (\S+), alignment at \((\S+), (\d+)-(\d+)\) skipped due to overflow", err)
        under = re.findall(r"Warning! Read (\S+), alignment at \((\S+), (\d+)-(\d+)\) skipped due to underflow", err)
        meta = {"args": satcase.ARGS, "n_overflow": len(over), "n_underflow": len(under),
                "first_overflow": list(over[0]) if over else None, "first_underflow": list(under[0]) if under else None,
                "lambda": [float(x) for x in re.findall(r"Background pileup value: ([0-9.]+)", err)],
                "peaks": int(re.search(r"Peaks identified: (\d+)", err).group(1))}
        meta["log_sha256"], meta["log_lines"] = sha(logf)
        meta["pile_sha256"], meta["pile_lines"] = sha(pile, skip_hash=True)
        h = hashlib.sha256()
        for a in over + under:                        # every skipped alignment, overflow first then underflow, each in file order
            h.update(("%s %s %s %s\n" % a).encode())
        meta["skipped_sha256"] = h.hexdigest()
        meta["verbose_sha256"] = hashlib.sha256(err.replace(sam, "T.sam").encode()).hexdigest()   # the whole -v text
        json.dump(meta, open(os.path.join(HERE, "sat_hot.json"), "w"), indent=1, sort_keys=True)
        print(meta)


if __name__ == "__main__":
    main()
