#!/usr/bin/env python
"""Instruction-count summary of the sm_100a kernels in libqeb_sm100.so (cuobjdump -sass): the Blackwell-specific mnemonics
that show which hardware paths a kernel uses - UTCHMMA / UTCIMMA.. (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus size and registers.
`python scripts/sass_summary.py > profiles/r2_sass_summary.md` (runs in the build container: no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "query-efficient-approx-to-improve-ocr_b200", "libqeb_sm100.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS",
        "LDGSTS", "REDG", "RED", "ATOMG", "ATOMS", "MUFU", "HMMA", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+).*?SHARED:(\d+)", res):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if cur and m:
            op = m.group(1)
            kernels[cur]["_n"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    kernels[cur][k] += 1
    names = demangle(list(kernels))
    used = [k for k in KEYS if any(c[k] for c in kernels.values())]
    print("# SASS summary of libqeb_sm100.so (sm_100a)\n")
    print("Built from HEAD with `__graft_entry__.build()`; `cuobjdump -sass`, counted per kernel. UTCHMMA = tcgen05.mma (kind::f16 / tf32), "
          "UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, "
          "SYNCS = mbarrier ops.\n")
    print("| kernel | instr | regs | " + " | ".join(used) + " |")
    print("|---|---:|---:|" + "---:|" * len(used))
    tot = collections.Counter()
    for k, c in kernels.items():
        short = re.sub(r"\(anonymous namespace\)::", "", names.get(k, k))
        short = re.sub(r"\(.*", "", short)[:90]
        r = regs.get(k, ("", ""))[0]
        print(f"| `{short}` | {c['_n']} | {r} | " + " | ".join(str(c[u]) if c[u] else "" for u in used) + " |")
        tot.update(c)
    print(f"| **total ({len(kernels)} kernels)** | {tot['_n']} | | " + " | ".join(str(tot[u]) for u in used) + " |")


if __name__ == "__main__":
    main()
