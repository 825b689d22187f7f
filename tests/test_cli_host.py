"""The host program's own C sources (genrich_b200/cli: option parsing, SAM / gz / BAM decode,
mate pairing, multimap weighting, interval transforms, narrowPeak / -f / -k writers, -v text)
checked WITHOUT a GPU: `make -C oracle cli_twin` links them against the CPU oracle instead of
libgenrich_cuda.so (oracle/cli_twin.c, test infrastructure).  Both sides of the C-ABI then use
glibc's libm like the reference, so the comparison with the files the unmodified reference wrote
(tests/golden) is BYTE FOR BYTE: every narrowPeak column, the sha256 of the whole -f and -k text,
the -v scalars.  The same program over the CUDA library is tests/test_gpu_cli.py."""
import hashlib
import os
import re
import subprocess

import pytest

import util
from cases import CASES, BY_NAME

TWIN = os.path.join(util.ORACLE_DIR, "_test", "genrich-b200-oracle")


@pytest.fixture(scope="module")
def twin():
    subprocess.check_call(["make", "-s", "-C", util.ORACLE_DIR, "cli_twin"])
    return TWIN


def _sha(path, skip_hash_lines=False):
    h = hashlib.sha256()
    n = 0
    with open(path, "rb") as f:
        for line in f:
            if skip_hash_lines and line.startswith(b"#"):
                continue
            h.update(line)
            n += 1
    return h.hexdigest(), n


def run_twin(twin, case, td, extra=(), files=None):
    tfiles, cfiles = files if files else util.write_case_sams(case, td)
    out, logf, pile = (os.path.join(td, x) for x in ("o.np", "o.f", "o.k"))
    cmd = [twin, "-t", ",".join(tfiles), "-o", out, "-f", logf, "-k", pile, "-v"] + case.ref_args() + list(extra)
    if any(c != "null" for c in cfiles):
        cmd += ["-c", ",".join(cfiles)]
    if case.bed:
        bedf = os.path.join(td, "x.bed")
        util.write_case_bed(case, bedf)
        cmd += ["-E", bedf]
    r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    return out, logf, pile, r.stderr


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_host_program_matches_reference_files(twin, case, tmp_path):
    out, logf, pile, err = run_twin(twin, case, str(tmp_path))
    meta, gold = util.golden(case)
    assert open(out).read().split("\n")[:-1] == gold                     # narrowPeak, every column
    assert _sha(logf) == (meta["log_sha256"], meta["log_lines"])         # the whole -f file
    assert _sha(pile, True) == (meta["pile_sha256"], meta["pile_lines"])  # the whole -k file (minus the path lines)
    lam = [float(x) for x in re.findall(r"Background pileup value: ([0-9.]+)", err)]
    fac = [float(x) for x in re.findall(r"Scaling factor for control pileup: ([0-9.]+)", err)]
    assert lam == meta["lambda"] and fac == meta["factor"]
    assert int(re.search(r"Genome length: (\d+)bp", err).group(1)) == meta["genome_len"]
    assert int(re.search(r"Peaks identified: (\d+)", err).group(1)) == meta["peaks"]
    assert int(re.search(r"Peaks identified: \d+ \((\d+)bp\)", err).group(1)) == meta["peak_bp"]
    assert ("All q-values are 1" in err) == meta["all_q_one"]
    assert len(re.findall(r"prevented from extending", err)) == meta["clamp_warnings"]
