"""SASS opcode summary of libou_b200.so (runs without a GPU): per kernel, the count of the instructions
that prove which hardware path it uses -- tcgen05 MMA (UTCHMMA / UTCQMMA ...), TMEM loads (LDTM), TMA
(UTMALDG / UTMASTG), tcgen05 commits / mbarriers (UTCBAR, SYNCS), legacy mma.sync (HMMA), distributed
shared-memory stores (STAS), plus the FP32 / memory instruction mix.
    python tools/sass_summary.py [> profiles/r2_sass_summary.txt]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "open_universe_b200" / "csrc" / "libou_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA",
        "STAS", "FFMA", "FFMA2", "FMUL", "FADD", "LDG", "STG", "LDS", "STS", "BAR", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            counts[name]["total"] += 1
            counts[name][op.split(".")[0]] += 1
    used = [k for k in KEYS if any(c[k] for c in counts.values())]
    print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}  (sm_100a); instruction counts per kernel")
    print(f"{'kernel':72s} " + " ".join(f"{k:>7s}" for k in used))
    for n, c in counts.items():
        print(f"{n[:72]:72s} " + " ".join(f"{c[k]:7d}" for k in used))


if __name__ == "__main__":
    sys.exit(main())
