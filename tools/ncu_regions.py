#!/usr/bin/env python
"""Attributes the executed warp instructions (and stall samples) of one kernel in an ncu report to the source lines of the
KERNEL'S OWN BODY, inlined callees included (ncu's source page credits inlined code to the callee's lines only). Joins ncu's
SASS page (per-address counters) with `nvdisasm -gi` of the same build's object file (per-address inline chains).

  python tools/ncu_regions.py gpurun_out/prof.ncu-rep ssao_cull althea_b200/build/frame_fast.o [top]
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, kern, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for k in range(1, len(rows)):  # several launches match: keep the first one's table
    if rows[k] and rows[k][0] == "Kernel Name":
        rows = rows[:k]
        break
name = rows[0][1]
hdr = rows[1]
iA, iS, iE, iN = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iW = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
sass = [(int(r[iA], 16), r[iS].strip(), int(r[iE] or 0), int(r[iN] or 0)) for r in rows[2:] if len(r) == len(hdr)]
base = sass[0][0]
# the mangled name of the kernel: find the function in the object whose demangled name matches
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    cubin = os.path.join(td, os.listdir(td)[0])
    dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
syms = re.findall(r"\.text\.(\S+)\s*:", dis)
dem = subprocess.run(["c++filt"] + syms, capture_output=True, text=True).stdout.splitlines()
target = None
for s, d in zip(syms, dem):
    if d.replace(" ", "") .startswith(name.replace(" ", "")[:40].replace("void", "").strip()) or name.split("(")[0].split()[-1].split("<")[0] in d and ("<(bool)0>" in name) == ("<false" in d or "Lb0" in s):
        target = s
        if name.split("(")[0].split()[-1].split("<")[0] in d:
            break
assert target, "kernel not found in " + obj
body = dis.split(".text.%s:" % target, 1)[1]
body = body.split("\n//---------------------", 1)[0]
outer, inner, chain = {}, {}, []
fresh = False
for line in body.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        if not fresh:
            chain, fresh = [], True
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", line)
    if m and chain:
        off = int(m.group(1), 16)
        outer[off], inner[off] = chain[-1], chain[0]
        fresh = False
src_cache = {}
def text(f, n):
    for root in ("althea_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:110] if n - 1 < len(src_cache[p]) else ""
    return ""
agg, tot, tots = {}, 0, 0
for addr, ins, ie, ns in sass:
    k = outer.get(addr - base, ("?", 0))
    a = agg.setdefault(k, [0, 0])
    a[0] += ie
    a[1] += ns
    tot += ie
    tots += ns
print("%s: %d warp instructions, %d samples, %d SASS instructions" % (name, tot, tots, len(sass)))
for k, (ie, ns) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100.0 * ie / max(tot, 1), 100.0 * ns / max(tots, 1), k[0], k[1], text(*k)))
