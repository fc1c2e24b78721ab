#!/usr/bin/env python
"""Warp instructions executed per CUDA source line (top lines): ncu_instr_lines.py <rep> <lib.so> <kernel> [top] [steps]"""
import csv, os, subprocess, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 100
maps = ncu_lines.line_map(lib, kern)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, body = rows[1], rows[2:]
best = min([k for k in maps if kern in k], key=lambda k: abs(len(maps[k]) - len(body)))
mp = maps[best]
base = int(body[0][0], 16)
ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = defaultdict(lambda: [0, 0, 0])
ti = 0
for r in body:
    off = int(r[0], 16) - base
    s = mp.get(off, (None, ""))[0]
    agg[s][0] += int(r[ii]); agg[s][1] += int(r[isamp]); agg[s][2] += 1
    ti += int(r[ii])
src = {}
print("SASS instructions in kernel: %d; executed %.0f warp-instr per SM per step" % (len(body), ti / 148.0 / steps))
for k, (a, b, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", k[0])
        if os.path.exists(path):
            if path not in src: src[path] = open(path).read().splitlines()
            if k[1] - 1 < len(src[path]): text = src[path][k[1] - 1].strip()
    print("%7.0f /SM/step %5.1f%%  sass %4d  %-24s %s" % (a / 148.0 / steps, 100.0 * a / ti, c, "%s:%d" % k if k else "?", text[:70]))
