#!/usr/bin/env python
"""Aggregates the ncu source page (per-SASS-instruction stall samples) by opcode class and by code
region, to see where a kernel's warps wait.  usage: python tools/ncu_source_regions.py rep.ncu-rep [nregions]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    nreg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = raw.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    by_op = collections.Counter()
    by_op_exec = collections.Counter()
    by_stall = collections.Counter()
    for r in rows:
        m = re.sub(r"^@!?U?P\w+\s+", "", r["Source"].strip())
        op = m.split()[0].split(".")[0] if m else "?"
        by_op[op] += int(r["# Samples"] or 0)
        by_op_exec[op] += int(r["Instructions Executed"] or 0)
        for c in stall_cols:
            by_stall[c] += int(r[c] or 0)
    print("total samples", tot)
    print("by stall:", [(k, v) for k, v in by_stall.most_common(10)])
    print("by opcode (samples, share, executed):")
    for op, v in by_op.most_common(14):
        print(f"   {op:10s} {v:8d} {v / tot:6.1%}  exec {by_op_exec[op]}")
    # regions of equal instruction count over the executed part
    live = [r for r in rows if int(r["Instructions Executed"] or 0) > 0]
    per = max(1, len(live) // nreg)
    print(f"regions of {per} instructions (executed only): start address, samples share, top stalls, first instr")
    for i in range(0, len(live), per):
        chunk = live[i:i + per]
        s = sum(int(r["# Samples"] or 0) for r in chunk)
        st = collections.Counter()
        for r in chunk:
            for c in stall_cols:
                st[c] += int(r[c] or 0)
        ex = sum(int(r["Instructions Executed"] or 0) for r in chunk) / len(chunk)
        top = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
        print(f"   {chunk[0]['Address'][-6:]} {s / tot:6.1%} exec/instr {ex:9.0f}  {top:55s} {chunk[0]['Source'][:40]}")


if __name__ == "__main__":
    main()
