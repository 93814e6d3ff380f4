#!/usr/bin/env python
"""Attribute an ncu source-page CSV (SASS level: `ncu -i x.ncu-rep --page source --csv`) to SOURCE LINES.

ncu's CSV export lists SASS instructions with their executed-instruction counts and stall samples but not the source
line; nvdisasm -g on the cubin of the same build (-lineinfo) lists the same instructions with `//## File ..., line N`
markers.  The two are joined by instruction offset inside the kernel.

    python tools/sass_lines.py <lib.so> <kernel substring> <source.csv[.gz]> [top N]
"""
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def disasm(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l]
    assert start, "kernel not found"
    i = start[0] + 1
    cur = ("?", 0)
    inl = ""
    table = []  # (offset, file, line, sass)
    while i < len(lines) and not lines[i].startswith("//--------------------- "):
        l = lines[i]
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            inl = m.group(3)
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            table.append((int(m.group(1), 16), cur[0], cur[1], m.group(2)))
        i += 1
    return table


def main():
    lib, kernel, path = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = disasm(lib, kernel)
    by_off = {o: (f, ln, s) for o, f, ln, s in table}
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hdr = rows[1]
    ia, ie, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(rows[2][ia], 16)
    per_line = defaultdict(lambda: [0, 0])
    tot_e = tot_s = 0
    miss = 0
    for r in rows[2:]:
        if len(r) <= max(ie, iss):
            continue
        off = int(r[ia], 16) - base
        e, s = int(r[ie] or 0), int(r[iss] or 0)
        tot_e += e
        tot_s += s
        if off in by_off:
            f, ln, _ = by_off[off]
        else:
            f, ln = "?", 0
            miss += 1
        per_line[(f, ln)][0] += e
        per_line[(f, ln)][1] += s
    print("total instructions executed %d, stall samples %d, unmatched SASS rows %d" % (tot_e, tot_s, miss))
    print("%-22s %6s %12s %6s %8s %6s" % ("file", "line", "inst", "%", "samples", "%"))
    for (f, ln), (e, s) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %6d %12d %5.1f%% %8d %5.1f%%" % (f, ln, e, 100.0 * e / max(1, tot_e), s, 100.0 * s / max(1, tot_s)))


if __name__ == "__main__":
    main()
