#!/usr/bin/env python
"""Ranks the CUDA source lines of one kernel in an ncu report by executed warp instructions (the source page, aggregated).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep ssao_cull [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
fname, hdr, out = None, None, []
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != "":
        i_ie, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            out.append((int(r[i_ie]), int(r[i_s] or 0), fname, r[0], r[1][:120]))
        except ValueError:
            pass
tot, tots = sum(o[0] for o in out), sum(o[1] for o in out)
print("total warp instructions %d, samples %d" % (tot, tots))
for o in sorted(out, reverse=True)[:top]:
    print("%5.2f%% inst %5.2f%% smp  %s:%s  %s" % (100.0 * o[0] / max(tot, 1), 100.0 * o[1] / max(tots, 1), o[2], o[3], o[4]))
