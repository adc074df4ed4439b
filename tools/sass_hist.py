#!/usr/bin/env python
"""Static SASS opcode histogram of the kernels in an object file.  usage: sass_hist.py obj.o [name-substring]"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if want not in name:
        continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
    print(name[:90], sum(ops.values()))
    print("  ", ops.most_common(30))
