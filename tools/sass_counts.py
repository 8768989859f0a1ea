#!/usr/bin/env python
"""Mnemonic counts and one excerpt per tensor-core / TMA mnemonic of the kernels in libqeft_b200.so (cuobjdump -sass).

    python tools/sass_counts.py [kernel-name-substring ...] > profiles/rNN_gemm_sass.txt

No GPU needed.  The mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor), SYNCS (mbarrier).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qeft_b200", "csrc", "libqeft_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "IMMA", "LDSM", "LDGSTS",
         "HFMA2", "LOP3", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "MEMBAR", "BAR"]


def main():
    want = sys.argv[1:] or ["gemm_w4_kernel", "gemm_w4_dx_kernel", "dow_kernel"]
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name, body = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            body[name] = []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            body[name].append(line)
    for name, lines in body.items():
        if not any(w in name for w in want):
            continue
        counts = collections.Counter()
        first = {}
        for ln in lines:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if not m:
                continue
            op = m.group(1)
            counts[op] += 1
            first.setdefault(op, ln.rstrip()[:110])
        short = re.sub(r"\(.*", "", name)
        print(f"# {short}: {len(lines)} instructions")
        print("  " + "  ".join(f"{w} {counts[w]}" for w in WATCH if counts[w]))
        for w in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP"):
            if w in first:
                print(first[w])
        print()


if __name__ == "__main__":
    main()
