"""profiles/<tag>_sass_mnemonics.txt: per kernel of libdsoft.so, the static counts of the SASS mnemonics that show what
the code runs on - tcgen05 MMAs (UTCHMMA), TMA loads / stores (UTMALDG / UTMASTG), bulk copies (UBLKCP), TMEM loads
(LDTM), tcgen05.commit (UTCBAR), packed fp32 (FFMA2 / FMUL2 / FADD2), MUFU, shuffles.  Needs cuobjdump only.

    python scripts/sass_mnemonics.py r04
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "refining-clip-via-dinov2-representations_b200", "libdsoft.so")
KEEP = re.compile(r"^(UTC|UTMA|UBLKCP|LDTM|STTM|SYNCS|UCGABAR|ELECT|MEMBAR|FFMA2|FMUL2|FADD2|MUFU|SHFL|REDUX|ATOM|RED)")


def main():
    tag = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Za-z0-9_.]+)", ln)
        if m and cur is not None and KEEP.match(m.group(1)):
            cur[m.group(1)] += 1
    out = os.path.join(ROOT, "profiles", f"{tag}_sass_mnemonics.txt")
    with open(out, "w") as f:
        f.write("cuobjdump -sass libdsoft.so (sm_100a): tensor / TMA / TMEM / cluster / packed-fp32 mnemonics per kernel "
                "(count of static instructions)\n")
        for fn, c in per.items():
            if "dsoft" not in fn or not c:
                continue
            f.write(f"\n{fn}\n")
            for k in sorted(c):
                f.write(f"    {k:<40s} {c[k]}\n")
    print(out)


if __name__ == "__main__":
    main()
