#!/usr/bin/env python
"""Joins an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm line info of the same
kernel and prints executed warp-instructions aggregated by source line and by opcode.

    python tools/ncu_by_line.py <src.csv> <cubin> <mangled-kernel-substring> [top]
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    src_csv, cubin, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = rows[hdr]
    ci = cols.index("Instructions Executed")
    si = cols.index("Source")
    st = cols.index("# Samples") if "# Samples" in cols else None
    insts = [(r[si].strip(), int(r[ci] or 0), int(r[st] or 0) if st is not None else 0) for r in rows[hdr + 1:] if len(r) > ci]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate the function
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
    lines = []
    cur = ("?", 0)
    stack = cur
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith("\t.section") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    if len(lines) != len(insts):
        print(f"warning: {len(lines)} disassembled vs {len(insts)} profiled instructions", file=sys.stderr)
    n = min(len(lines), len(insts))
    by_line = collections.Counter()
    by_line_s = collections.Counter()
    by_op = collections.Counter()
    total = 0
    for (f, ln), (sass, cnt, smp) in zip(lines[:n], insts[:n]):
        by_line[(f, ln)] += cnt
        by_line_s[(f, ln)] += smp
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        by_op[op.split(".")[0]] += cnt
        total += cnt
    tot_s = sum(by_line_s.values()) or 1
    print(f"total warp-instructions {total:.3e}")
    print("--- by source line (share of instructions, share of stall samples)")
    for (f, ln), c in by_line.most_common(top):
        print(f"{100 * c / total:6.2f}%  {100 * by_line_s[(f, ln)] / tot_s:6.2f}%  {f}:{ln}")
    print("--- by opcode")
    for op, c in by_op.most_common(30):
        print(f"{100 * c / total:6.2f}%  {op}")


if __name__ == "__main__":
    main()
