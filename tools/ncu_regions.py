#!/usr/bin/env python
"""Instruction and stall-sample share per kernel phase (regions are delimited by PROF_MARK lines).
usage: ncu_regions.py <report.ncu-rep> <lib.so> <source.cu> [kernel-substring] [steps]"""
import csv, re, subprocess, sys, os
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

rep, lib, srcfile = sys.argv[1:4]
kern = sys.argv[4] if len(sys.argv) > 4 else "sv_fast_kernel"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 100
src = open(srcfile).read().splitlines()
marks = [(i + 1, re.search(r"PROF_MARK\((\d+)\);\s*//\s*(.*)", l)) for i, l in enumerate(src)]
marks = [(ln, m.group(2)) for ln, m in marks if m]
kstart = next(i + 1 for i, l in enumerate(src) if "__global__" in l and kern in l)
maps = ncu_lines.line_map(lib, kern)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, body = rows[1], rows[2:]
best = min([k for k in maps if kern in k], key=lambda k: abs(len(maps[k]) - len(body)))
mp = maps[best]
base = int(body[0][0], 16)
ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
base_name = os.path.basename(srcfile)


def region(srcline):
    if srcline is None:
        return "?"
    f, l = srcline
    if f != base_name:
        return "inlined:" + f
    if l < kstart:
        return "device helpers"
    prev = "prologue/init"
    for ln, name in marks:
        if l <= ln:
            return "-> " + name
        prev = name
    return "after last mark"


agg = defaultdict(lambda: [0, 0])
ti = ts = 0
for r in body:
    off = int(r[0], 16) - base
    s = mp.get(off, (None, ""))[0]
    reg = region(s)
    agg[reg][0] += int(r[ii])
    agg[reg][1] += int(r[isamp])
    ti += int(r[ii])
    ts += int(r[isamp])
print("total warp instructions %.3e (%.0f per SM per step), samples %d" % (ti, ti / 148.0 / steps, ts))
for k, (a, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s instr %5.1f%% (%7.0f /SM/step)   samples %5.1f%%" % (k[:50], 100.0 * a / ti, a / 148.0 / steps, 100.0 * b / ts))
