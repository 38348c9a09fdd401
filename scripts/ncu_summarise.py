#!/usr/bin/env python
"""Turn the raw ncu outputs that come back in gpurun_out/ into the small, tracked summaries under profiles/.

    python scripts/ncu_summarise.py launches <launches.csv> <out_summary.csv> "<command that was profiled>"
    python scripts/ncu_summarise.py full <prof.ncu-rep> <out.json> <workload>   (also updates profiles/ncu_traffic.json)

`launches`: per-kernel launch count, mean device time and SHARE of the step (the per-launch times of
an ncu pass are cold-cache and serialised, so only the shares are comparable with bench.py's events).
`full`: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), DRAM throughput, warps
active, registers of every profiled kernel; bench.py reads profiles/ncu_traffic.json for `roofline.traffic`.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0]


def launches(src, dst, cmd):
    rows = [l for l in open(src) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    acc = {}
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        if k.startswith("k_generate_mask") or not k.startswith("k_"):
            continue  # input synthesis / torch fill kernels are outside the step
        v = float(r["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r["Metric Unit"]]
        n, s = acc.get(k, (0, 0.0))
        acc[k] = (n + 1, s + v)
    total = sum(s for _, s in acc.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary; command: %s\n" % cmd)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_ms\n")
        f.write("kernel,launches,avg_us,share_of_step_pct\n")
        for k, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.1f\n" % (k, n, s / n, 100.0 * s / total))
    print(open(dst).read())


WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def full(rep, dst, workload):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for m, key in WANT.items():
            if m in hdr:
                i = hdr.index(m)
                v = float(r[i].replace(",", ""))
                d[key] = v * SCALE.get(units[i], 1.0) if key in ("dram_read", "dram_write", "duration") else v
        d["duration_us"] = d.pop("duration")
        d["dram_bytes"] = d["dram_read"] + d["dram_write"]
        d["dram_gbs"] = d["dram_bytes"] / d["duration_us"] / 1e3
        res.append(d)
    with open(dst, "w") as f:
        json.dump({"workload": workload, "source": os.path.basename(rep),
                   "how": "ncu --set full --clock-control none --import-source on (one launch each, ~40 replays)",
                   "kernels": res}, f, indent=1)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t.setdefault(workload, {})
    for d in res:
        t[workload][d["kernel"].split("<")[0]] = int(d["dram_bytes"])
    json.dump(t, open(tpath, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
