#!/usr/bin/env python
"""Static SASS evidence per kernel of liblrds_b200.so: tcgen05.mma (UTC*MMA), tcgen05.ld / st (LDTM / STTM), bulk TMA
(UBLKCP), mbarrier (SYNCS), packed fp32x2 (FFMA2), registers.   python tools/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sde_sampler_lrds_b200", "csrc", "liblrds_b200.so")
PATTERNS = [("UTCxMMA", r"\bUTC[A-Z]*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UBLKCP", r"\bUBLKCP\b"),
            ("SYNCS", r"\bSYNCS\b"), ("FFMA2", r"\bFFMA2\b")]


def main():
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
        print(f"{'kernel':<100} " + " ".join(f"{n:>7}" for n, _ in PATTERNS) + "  regs")
        for cubin in sorted(f for f in os.listdir(tmp) if f.startswith("lrds_tc_") and f.endswith(".cubin")):
            path = os.path.join(tmp, cubin)
            sass = subprocess.run(["nvdisasm", "-c", path], capture_output=True, text=True).stdout.splitlines()
            res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
            regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
            counts, cur = collections.defaultdict(collections.Counter), None
            for line in sass:
                if line.startswith(".text."):
                    cur = line[len(".text."):].rstrip(":")
                elif cur:
                    for name, pat in PATTERNS:
                        if re.search(pat, line):
                            counts[cur][name] += 1
            for k in sorted(counts):
                if not counts[k]["UTCxMMA"]:
                    continue
                dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
                dem = re.sub(r"\(lrds::RolloutArgs.*", "", dem).replace("void lrds::", "")
                print(f"{dem:<100} " + " ".join(f"{counts[k][n]:>7}" for n, _ in PATTERNS) + f"  {regs.get(k, '?'):>4}")


if __name__ == "__main__":
    main()
