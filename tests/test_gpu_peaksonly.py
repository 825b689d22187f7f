"""-P on the device (gr_load_pvalues + K8) through the host program."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_peaks_only_on_device(tmp_path):
    """-P on the device: the host program parses a -f log it wrote itself and K8 runs on the loaded
    -log(p) / -log(q) values (gr_load_pvalues); narrowPeak files equal the unmodified reference's
    (tests/golden/ponly_*)."""
    import json
    import util
    from cases import BY_NAME
    cli = os.path.join(ROOT, "genrich_b200", "bin", "genrich-b200")
    golden = os.path.join(ROOT, "tests", "golden")
    logs = {}
    for name in sorted(f[:-5] for f in os.listdir(golden) if f.startswith("ponly_") and f.endswith(".json")):
        meta = json.load(open(os.path.join(golden, name + ".json")))
        cname = meta["case"]
        case = BY_NAME[cname]
        if cname not in logs:
            td = str(tmp_path / cname)
            os.makedirs(td)
            tfiles, cfiles = util.write_case_sams(case, td)
            logf = os.path.join(td, "o.f")
            cmd = [cli, "-t", ",".join(tfiles), "-o", os.path.join(td, "o.np"), "-f", logf] + case.ref_args()
            if any(c != "null" for c in cfiles):
                cmd += ["-c", ",".join(cfiles)]
            if case.bed:
                bedf = os.path.join(td, "x.bed")
                util.write_case_bed(case, bedf)
                cmd += ["-E", bedf]
            subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL, timeout=120)
            logs[cname] = logf
        out = str(tmp_path / (name + ".np"))
        cmd = [cli, "-P", "-f", logs[cname], "-o", out] + meta["args"]
        if meta["bed_case"]:
            bedf = str(tmp_path / (name + ".bed"))
            util.write_case_bed(BY_NAME[meta["bed_case"]], bedf)
            cmd += ["-E", bedf]
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True, timeout=120)
        assert r.returncode == 0, (name, r.stderr)
        got = open(out).read().split("\n")[:-1]
        want = open(os.path.join(golden, name + ".narrowPeak")).read().split("\n")[:-1]
        assert len(got) == len(want) == meta["peaks"], name
        for g, w in zip(got, want):
            gf, wf = g.split("\t"), w.split("\t")
            assert gf[:4] == wf[:4] and gf[9] == wf[9], (name, g, w)          # name, start, end, peak_N, summit
            for i in (6, 7, 8):      # the device's own -f text can differ from the reference's in the 6th decimal of p / q;
                                     # the AUC (column 6) sums such differences over up to ~1000 bp
                assert abs(float(gf[i]) - float(wf[i])) <= 1e-4 + (1e-5 if i == 6 else 2e-6) * abs(float(wf[i])), (name, g, w)
