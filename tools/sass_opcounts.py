#!/usr/bin/env python
"""Mnemonic histogram per kernel of the built library's SASS (cuobjdump -sass), the tracked evidence that the hot kernels
are tcgen05 / TMEM / TMA code:  python tools/sass_opcounts.py > profiles/r02_sass_opcounts.txt
Columns: total instructions, then the Blackwell tensor-path mnemonics of B200_PROFILING.md (UTCHMMA = tcgen05.mma, LDTM /
STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit) and the legacy HMMA (mma.sync)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "instarevive_b200" / "csrc" / "libinstarevive_b200.so"
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "MUFU", "SYNCS", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    funcs = OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(next(it), Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            op = m.group(1)
            if op in KEYS:
                cur[op] += 1
                if op in ("UTCHMMA", "UTMALDG", "UTCBAR") and ".2CTA" in m.group(2):
                    cur[op + ".2CTA"] += 1
    total = Counter()
    print(f"# {LIB.name}: {len(funcs)} kernels; cuobjdump -sass mnemonic counts (static instructions, not executed counts)")
    cols = KEYS + ["UTCHMMA.2CTA", "UTMALDG.2CTA"]
    print("kernel".ljust(72) + "total".rjust(8) + "".join(c.rjust(14) for c in cols))
    for name, c in sorted(funcs.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["_total"]):
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("ir::", "")
        print(short[:71].ljust(72) + str(c["_total"]).rjust(8) + "".join(str(c[k]).rjust(14) for k in cols))
        total.update(c)
    print("TOTAL".ljust(72) + str(total["_total"]).rjust(8) + "".join(str(total[k]).rjust(14) for k in cols))


if __name__ == "__main__":
    sys.exit(main())
