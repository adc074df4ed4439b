#!/usr/bin/env python
"""Single-warp issue timeline of a kernel's SASS (the event-step model of /opt/skills/guides/B300_MICROARCH.md):
walks the instructions of the main loop in order, honouring each instruction's stall count and scoreboard wait
mask, with nominal latencies for the variable-latency classes, and prints per source region how many cycles a
warp running ALONE would spend there, how many FP32-pipe cycles it asks for, and the issue slots it uses.
Where a region's alone-time is far above its pipe demand, a warp cannot feed the pipe there by itself and the
other warps of the scheduler have to.

usage: python tools/sass_timeline.py lib.so kernel-substring [--from ADDR --to ADDR] [--marks ADDR,ADDR,...]
"""
import argparse
import re
import subprocess

LAT = {"LDS": 30, "LDG": 600, "SHFL": 26, "MUFU": 22, "BAR": 40, "STS": 6, "STG": 6, "S2R": 20, "LDC": 40, "LDCU": 40,
       "F2I": 12, "I2F": 12, "SYNCS": 30, "UBLKCP": 30, "ATOM": 300, "RED": 10, "LDL": 30, "STL": 6, "WARPSYNC": 4,
       "ELECT": 10, "R2UR": 12, "S2UR": 20, "BMOV": 12, "REDUX": 20, "VOTE": 12, "MATCH": 20}
FMA2 = ("FFMA2", "FADD2", "FMUL2")
FMA1 = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2")


def parse(lib, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    out, on = [], False
    i = 0
    while i < len(txt):
        line = txt[i]
        if "Function :" in line:
            on = kernel in line
        elif on:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", line)
            if m and i + 1 < len(txt):
                hi = re.search(r"/\* 0x([0-9a-f]{16}) \*/", txt[i + 1])
                if hi:
                    h = int(hi.group(1), 16)
                    text = m.group(2).strip()
                    op = re.sub(r"^@!?U?P\w+\s+", "", text).split()[0]
                    out.append({"addr": int(m.group(1), 16), "text": text, "op": op.split(".")[0], "full": op,
                                "stall": (h >> 41) & 0xF, "yield": (h >> 45) & 1, "wbar": (h >> 46) & 7,
                                "rbar": (h >> 49) & 7, "wait": (h >> 52) & 0x3F})
                    i += 1
        i += 1
    return out


def simulate(ins):
    t, sb = 0, [0] * 6
    rows = []
    for x in ins:
        arm = max([sb[s] for s in range(6) if x["wait"] >> s & 1], default=0)
        start = max(t, arm)
        lat = LAT.get(x["op"], 12)
        if x["wbar"] < 6:
            sb[x["wbar"]] = max(sb[x["wbar"]], start + lat)
        if x["rbar"] < 6:
            sb[x["rbar"]] = max(sb[x["rbar"]], start + 6)
        rows.append((x, start, start - t))
        t = start + max(1, x["stall"])
    return rows, t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("kernel")
    ap.add_argument("--lo", type=lambda s: int(s, 16), default=None)
    ap.add_argument("--hi", type=lambda s: int(s, 16), default=None)
    ap.add_argument("--marks", default="")
    ap.add_argument("--dump", action="store_true")
    a = ap.parse_args()
    ins = parse(a.lib, a.kernel)
    if not ins:
        raise SystemExit("kernel not found")
    if a.lo is None:
        # main loop: the last backward branch with the longest span
        best = None
        for x in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", x["text"])
            if m and int(m.group(1), 16) < x["addr"]:
                span = x["addr"] - int(m.group(1), 16)
                if best is None or span > best[0]:
                    best = (span, int(m.group(1), 16), x["addr"])
        a.lo, a.hi = best[1], best[2]
    body, skip_to = [], None
    for x in ins:
        if not (a.lo <= x["addr"] <= a.hi) or (skip_to is not None and x["addr"] < skip_to):
            continue
        skip_to = None
        body.append(x)
        m = re.match(r"BRA\s+0x([0-9a-f]+)", x["text"])     # unpredicated forward branch: the hot path jumps over cold blocks
        if m and int(m.group(1), 16) > x["addr"]:
            skip_to = int(m.group(1), 16)
    rows, total = simulate(body)
    print(f"loop {a.lo:#x}..{a.hi:#x}: {len(body)} instructions, single-warp time {total} cycles (fall-through path incl. cold blocks)")
    marks = sorted(int(m, 16) for m in a.marks.split(",") if m) or [body[0]["addr"] + k * ((a.hi - a.lo) // 16 // 16 * 16) for k in range(16)]
    marks = [m for m in marks if a.lo <= m <= a.hi]
    if not marks or marks[0] != a.lo:
        marks = [a.lo] + marks
    marks.append(a.hi + 16)
    print("region            instr  alone  fma_cyc  sb_wait  stall>2  top ops")
    for lo, hi in zip(marks[:-1], marks[1:]):
        seg = [(x, s, w) for (x, s, w) in rows if lo <= x["addr"] < hi]
        if not seg:
            continue
        dur = (seg[-1][1] + max(1, seg[-1][0]["stall"])) - seg[0][1] + seg[0][2]
        fma = sum(2 if x["op"] in FMA2 else 1 if x["op"] in FMA1 else 0 for x, _, _ in seg)
        sbw = sum(w for _, _, w in seg)
        st = sum(max(0, x["stall"] - 2) for x, _, _ in seg if x["op"] in FMA2) + sum(max(0, x["stall"] - 1) for x, _, _ in seg if x["op"] not in FMA2)
        ops = {}
        for x, _, _ in seg:
            ops[x["op"]] = ops.get(x["op"], 0) + 1
        top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
        print(f"{lo:#7x}..{hi:#7x} {len(seg):6d} {dur:6d} {fma:8d} {sbw:8d} {st:8d}  {top}")
    if a.dump:
        for x, s, w in rows:
            print(f"{x['addr']:#6x} t={s:6d} w={w:4d} st={x['stall']:2d} y={x['yield']} wb={x['wbar']} rb={x['rbar']} wm={x['wait']:02x}  {x['text']}")


if __name__ == "__main__":
    main()
