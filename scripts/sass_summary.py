#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libb200vae.so (runs without a GPU: cuobjdump disassembles the sm_100a cubin).

    python scripts/sass_summary.py > profiles/r02_sass_opcodes.txt

Counts, per kernel, the opcodes that identify Blackwell-native code: UTCHMMA (tcgen05.mma; `.2CTA` = cta_group::2), LDTM /
STTM (tcgen05.ld / st: TMEM <-> registers), UTMALDG (TMA tensor loads), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
FFMA2 / FMUL2 (packed f32x2 arithmetic), MUFU, plus the total instruction count."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vae_song_b200", "libb200vae.so")
WATCH = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.2CTA", "UTCBAR", "SYNCS", "FFMA2", "FMUL2", "FFMA", "MUFU",
         "LDGSTS", "ATOM", "RED")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            cur["total"] += 1
            cur[op] += 1
            if op in ("UTCHMMA", "UTMALDG") and ".2CTA" in mods:
                cur[op + ".2CTA"] += 1
    names = demangle(list(per))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels, sm_100a SASS opcode counts (scripts/sass_summary.py)")
    print("# kernel | total | " + " | ".join(WATCH))
    tot = collections.Counter()
    for k, c in per.items():
        short = re.sub(r"\(.*", "", names.get(k, k)).replace("b200vae::", "")
        short = re.sub(r"^void ", "", short)
        print(f"{short} | {c['total']} | " + " | ".join(str(c[w]) for w in WATCH))
        tot.update(c)
    print(f"ALL | {tot['total']} | " + " | ".join(str(tot[w]) for w in WATCH))


if __name__ == "__main__":
    sys.exit(main())
