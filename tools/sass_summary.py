#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the built library: which kernels use the 5th-gen tensor cores (UTCHMMA = tcgen05.mma,
.2CTA = cta_group::2), TMEM loads/stores (LDTM / STTM), TMA (UTMALDG tiled / IM2COL, UTMASTG, UBLKCP = cp.async.bulk) and
which still use the warp-level MMA (HMMA).  Runs on the CPU-only build box (cuobjdump reads the cubin inside the .so).

    python tools/sass_summary.py > profiles/rNN_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "imageenhancement_mp_b200", "libimgenh_b200.so")
PATTERNS = [("UTCHMMA.2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("UTCHMMA", r"\bUTC[A-Z]*MMA\b(?!\.2CTA)"), ("UTCBAR", r"\bUTCBAR"),
            ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG.IM2COL", r"\bUTMALDG\.[0-9]D\.IM2COL"),
            ("UTMALDG", r"\bUTMALDG\.[0-9]D(?!\.IM2COL)"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
            ("UTMAPF", r"\bUTMAPF|\bUTMACCTL"), ("HMMA", r"\bHMMA"), ("FFMA2", r"\bFFMA2"), ("SHFL", r"\bSHFL"),
            ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instructions"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    cols = ["instructions"] + [n for n, _ in PATTERNS]
    print("# cuobjdump -sass imageenhancement_mp_b200/libimgenh_b200.so : opcode counts per kernel (sm_100a)")
    print("# " + "  ".join(cols))
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        tot.update(c)
        print(f"{k[:70]:70s} " + " ".join(f"{n}={c[n]}" for n in cols if c[n]))
    print(f"{'TOTAL':70s} " + " ".join(f"{n}={tot[n]}" for n in cols if tot[n]))


if __name__ == "__main__":
    sys.exit(main())
