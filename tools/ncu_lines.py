#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS view) per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
Needs the library compiled with -lineinfo; uses cuobjdump/nvdisasm for the offset->line map.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_map(lib, kern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        if kern not in txt:
            continue
        maps, cur, name = {}, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                name = m.group(1)
                maps[name] = {}
                cur = None
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
            if m and name:
                maps[name][int(m.group(1), 16)] = (cur, m.group(2).strip())
        return maps
    return {}


def main():
    rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    maps = line_map(lib, kern)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # several launches may match: keep the first block
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if len(starts) > 1:
        rows = rows[:starts[1]]
    kname = rows[0][1]
    mangled = [k for k in maps if kern in k]
    # pick the map whose instruction count matches best
    hdr = rows[1]
    body = rows[2:]
    best = min(mangled, key=lambda k: abs(len(maps[k]) - len(body)))
    mp = maps[best]
    base = int(body[0][0], 16)
    isamp = hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    per_line = defaultdict(lambda: defaultdict(int))
    tot = 0
    for r in body:
        off = int(r[0], 16) - base
        src = mp.get(off, (None, ""))[0]
        s = int(r[isamp])
        tot += s
        per_line[src]["samples"] += s
        for i in stall_cols:
            per_line[src][hdr[i]] += int(r[i])
    print("kernel:", kname, " total samples:", tot)
    srcs = {}
    for (src, d) in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        reasons = sorted(((v, k) for k, v in d.items() if k != "samples"), reverse=True)[:3]
        text = ""
        if src:
            path = None
            for root in (os.path.dirname(os.path.abspath(lib)) + "/../csrc",
                         os.path.dirname(os.path.abspath(lib)) + "/pmmh-qn_b200/csrc"):
                p = os.path.join(root, src[0])
                if os.path.exists(p):
                    path = p
            if path:
                if path not in srcs:
                    srcs[path] = open(path).read().splitlines()
                if src[1] - 1 < len(srcs[path]):
                    text = srcs[path][src[1] - 1].strip()
        print("%6.2f%%  %-22s %-60s  %s" % (100.0 * d["samples"] / max(tot, 1),
                                           "%s:%d" % src if src else "?", text[:60],
                                           " ".join("%s=%.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(d["samples"], 1)) for v, k in reasons)))


if __name__ == "__main__":
    main()
