#!/usr/bin/env python
"""DRAM bytes and device time per kernel of ONE step of bench.py, from an ncu pass:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/dram_<workload>.csv python bench.py --profile --workload <workload>
    python tools/ncu_dram_by_stage.py gpurun_out/dram_<workload>.csv gpurun_out/profile_meta.json \
        profiles/r02_dram_by_stage.json [profiles/r02_dram_<workload>.txt]

`bench.py --profile` runs warm-up steps and ONE step and writes how many kernels a step launches
(profile_meta.json); the last that-many launches of the log are the step.  The JSON (one entry per
"<workload>@<ranks>") is what bench.py loads for roofline.traffic / roofline.step -- by kernel name,
so a renamed kernel fails loudly there instead of quoting stale bytes."""
import collections
import csv
import json
import os
import sys


def main():
    log, meta_p, out_p = sys.argv[1:4]
    txt_p = sys.argv[4] if len(sys.argv) > 4 else None
    meta = json.load(open(meta_p))
    rows = list(csv.reader(l for l in open(log) if l.startswith('"')))
    h = rows[0]
    iid, ik, im, iu, iv = (h.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), {"kernel": r[ik].split("(")[0].replace("void ", "").split("<")[0]})
        v = float(r[iv].replace(",", ""))
        unit = r[iu].lower()
        if r[im].startswith("dram__bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1)
        elif r[im].startswith("gpu__time"):
            v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        d[r[im]] = v
    ids = sorted(launches)
    n = int(meta["launches_per_step"])
    if len(ids) < n:
        raise SystemExit("log holds %d launches, a step has %d" % (len(ids), n))
    step = [launches[i] for i in ids[-n:]]
    ks = collections.OrderedDict()
    for d in step:
        a = ks.setdefault(d["kernel"], {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "time_ns": 0.0})
        a["launches"] += 1
        a["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
        a["time_ns"] += d.get("gpu__time_duration.sum", 0.0)
    key = "%s@%d" % (meta["workload"], meta.get("world", 1))
    table = json.load(open(out_p)) if os.path.exists(out_p) else {}
    tot_b = sum(v["dram_read"] + v["dram_write"] for v in ks.values())
    tot_t = sum(v["time_ns"] for v in ks.values())
    table[key] = {"command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                             "--clock-control none python bench.py --profile --workload %s" % meta["workload"],
                  "launches_per_step": n, "step_dram_bytes": tot_b, "step_kernel_time_ns_under_ncu": tot_t, "kernels": ks}
    json.dump(table, open(out_p, "w"), indent=1)
    lines = ["%s: one step = %d launches, %.3f GB through DRAM, %.3f ms of kernels (serialised, cold cache: shares, not absolutes)"
             % (key, n, tot_b / 1e9, tot_t / 1e6),
             "%-22s %4s %10s %10s %10s %8s %9s" % ("kernel", "n", "read GB", "write GB", "ms", "share", "GB/s")]
    for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["time_ns"]):
        b = v["dram_read"] + v["dram_write"]
        lines.append("%-22s %4d %10.3f %10.3f %10.3f %7.1f%% %9.0f" % (
            k, v["launches"], v["dram_read"] / 1e9, v["dram_write"] / 1e9, v["time_ns"] / 1e6,
            100 * v["time_ns"] / max(tot_t, 1), b / max(v["time_ns"], 1)))
    print("\n".join(lines))
    if txt_p:
        open(txt_p, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
