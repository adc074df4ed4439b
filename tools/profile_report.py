#!/usr/bin/env python
"""Collects the round's measurement artefacts from gpurun_out/ into profiles/ (tracked): bench lines, the ncu
launch list with per-kernel shares, the ncu --set full summary of the headline kernel, DRAM traffic per frame.
usage: python tools/profile_report.py r01 --bench gpurun_out/bench.json --ref gpurun_out/bench_ref.json
           --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep --frames-per-launch 109824 [--configs x.jsonl]"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--bench")
ap.add_argument("--ref")
ap.add_argument("--launches")
ap.add_argument("--rep")
ap.add_argument("--frames-per-launch", type=int, default=0)
ap.add_argument("--configs")
ap.add_argument("--scaling", nargs="*")
ap.add_argument("--note", default="")
a = ap.parse_args()
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def last_json_line(path):
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l.strip().startswith("{")]
    return json.loads(lines[-1])


if a.bench:
    json.dump(last_json_line(a.bench), open(os.path.join(P, f"{a.tag}_bench.json"), "w"), indent=1)
if a.ref:
    json.dump(last_json_line(a.ref), open(os.path.join(P, f"{a.tag}_bench_reference.json"), "w"), indent=1)
if a.configs:
    shutil.copy(a.configs, os.path.join(P, f"{a.tag}_configs.jsonl"))
if a.scaling:
    rows = [last_json_line(p) for p in a.scaling]
    keep = [{k: r.get(k) for k in ("n_gpus", "value", "unit", "ms_per_step", "e2e", "clocks", "kernel")} for r in rows]
    json.dump(keep, open(os.path.join(P, f"{a.tag}_scaling.json"), "w"), indent=1)
if a.launches:
    shutil.copy(a.launches, os.path.join(P, f"{a.tag}_launches.csv"))
    txt = open(a.launches).read()
    start = txt.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        tot[r["Kernel Name"][:60]] += us
        cnt[r["Kernel Name"][:60]] += 1
    s = sum(tot.values())
    with open(os.path.join(P, f"{a.tag}_launch_shares.txt"), "w") as f:
        f.write("kernel, launches, total us, share  (ncu --metrics gpu__time_duration.sum --clock-control none over a short "
                "bench.py run; cold-cache serialised times: compare shares, not absolutes)\n")
        if a.note:
            f.write(a.note + "\n")
        for k, v in tot.most_common():
            f.write(f"{k}, {cnt[k]}, {v:.1f}, {v / s:.1%}\n")
if a.rep:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), a.rep], capture_output=True, text=True).stdout
    out += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_regions.py"), a.rep, "24"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[0], rows[2]
    name = data[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("sg::", "")
    short = name.split("<")[0]
    open(os.path.join(P, f"{a.tag}_{short}_ncu_summary.txt"), "w").write(out)
    rd = float(data[hdr.index("dram__bytes_read.sum")].replace(",", ""))
    wr = float(data[hdr.index("dram__bytes_write.sum")].replace(",", ""))
    ur, uw = rows[1][hdr.index("dram__bytes_read.sum")], rows[1][hdr.index("dram__bytes_write.sum")]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd, wr = rd * scale[ur], wr * scale[uw]
    if a.frames_per_launch:
        json.dump({"kernel": name, "source": f"profiles/{a.tag}_{short}_ncu_summary.txt (ncu --set full, one launch)",
                   "frames_per_launch": a.frames_per_launch, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "dram_bytes_per_frame": (rd + wr) / a.frames_per_launch, "algorithmic_bytes_per_frame": 3072},
                  open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
print("profiles/ updated:", sorted(os.listdir(P)))
