#!/usr/bin/env python
"""Static evidence about the built library, no GPU needed: per-kernel registers / stack / shared memory
(`cuobjdump --dump-resource-usage`) and the SASS mnemonics that show what each kernel's memory path is (bulk-TMA copies
and mbarriers, 128-bit loads, local-memory spills). Usage: python tools/static_report.py > profiles/<round>_static.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "althea_b200", "lib", "libalthea_cuda.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o) for o in out]


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            c = counts[cur]
            c["total"] += 1
            for key, pat in (("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("LDG.128", r"^LDG\.E.*\.128"),
                             ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STL", r"^STL"), ("LDL", r"^LDL"),
                             ("ATOMG", r"^(ATOMG|RED)"), ("MUFU", r"^MUFU"), ("FFMA", r"^FFMA")):
                if re.match(pat, op):
                    c[key] += 1
    names = sorted(usage)
    pretty = dict(zip(names, demangle(names)))
    print("# Static report of althea_b200/lib/libalthea_cuda.so (sm_100a)\n")
    print("`cuobjdump --dump-resource-usage` and instruction counts from `cuobjdump -sass`; produced by "
          "`tools/static_report.py`. STL/LDL are local-memory (spill or stack array) instructions; UBLKCP + SYNCS are the "
          "bulk-TMA copy and its mbarrier.\n")
    cols = ["total", "FFMA", "MUFU", "LDG", "LDG.128", "STG", "LDS", "UBLKCP", "SYNCS", "ATOMG", "STL", "LDL"]
    print("| kernel | regs | stack B | smem B | " + " | ".join(cols) + " |")
    print("|---|---|---|---|" + "---|" * len(cols))
    for n in names:
        reg, stack, shared, _local = usage[n]
        c = counts.get(n, {})
        print(f"| `{pretty[n]}` | {reg} | {stack} | {shared} | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
