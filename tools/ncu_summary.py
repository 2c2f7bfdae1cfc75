#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep, `ncu --set full`) or a launch list CSV into the text summaries committed
under profiles/. Runs on the CPU box (ncu -i needs no GPU).

  python tools/ncu_summary.py report  gpurun_out/prof.ncu-rep  profiles/r1_frame_full.md
  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe LSU wavefronts % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests (warp-level)"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by L1"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("smsp__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
]
STALLS_PREFIX = "smsp__average_warp_latency_issue_stalled_"
STALLS_ALT = "smsp__average_warps_issue_stalled_"


def report(path: str, out: str):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full --clock-control none: %s" % path.split("/")[-1], "",
             "One launch per section; numbers are per launch (cold-cache, serialised under the profiler: compare shares and",
             "utilisations, not absolute times, with the CUDA-event timings in bench.py's JSON line).", ""]
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        lines.append("## %s  (id %s)" % (name, r[col["ID"]]))
        for key, label in METRICS:
            if key in col and r[col[key]] != "":
                lines.append("- %-42s %s %s   `%s`" % (label, r[col[key]], units[col[key]], key))
        stalls = []
        for h, i in col.items():
            if h.startswith(STALLS_ALT) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALLS_ALT):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            lines.append("- top warp stall reasons (stalled warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


def launches(path: str, out: str):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        k = r[ik].split("(")[0]
        agg.setdefault(k, []).append((float(r[iv].replace(",", "")) / 1e6, r[ig], r[ib]))
    total = sum(v[0] for vs in agg.values() for v in vs)
    lines = ["# ncu launch list (gpu__time_duration.sum, --clock-control none): %s" % path.split("/")[-1], "",
             "Cold-cache, serialised launches: the SHARE column is what must agree with bench.py's per-kernel CUDA-event timings.", "",
             "| kernel | launches | total ms | share | per-launch ms (first 6) | grid | block |", "|---|---|---|---|---|---|---|"]
    for k, vs in agg.items():
        t = sum(v[0] for v in vs)
        lines.append("| %s | %d | %.3f | %.1f %% | %s | %s | %s |" % (k, len(vs), t, 100 * t / total, " ".join("%.3f" % v[0] for v in vs[:6]), vs[0][1], vs[0][2]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


def traffic(path: str, out: str):
    """profiles/ncu_traffic.json: per kernel, DRAM bytes per launch + the busiest unit (what bench.py's roofline.traffic quotes)."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rec = json.load(open(out)) if os.path.exists(out) else {}
    busy = {"L1 data pipe (LSU wavefronts)": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "L2 (lts throughput)": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "DRAM": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "instruction issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "FMA pipe": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "XU pipe": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("_kernel", "").split("::")[-1]
        name = name.replace("void ", "").split("<")[0]  # template instantiations (ssao<0>, glossy_convolve_staged<32, 32>) share a row
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
        util = {k: float(r[col[m]].replace(",", "")) for k, m in busy.items() if m in col and r[col[m]] != ""}
        top = max(util, key=util.get)
        rec[name] = {"dram_bytes_per_launch": rd + wr, "source": "profiles/" + os.path.basename(path).replace(".ncu-rep", "") + " (ncu --set full)",
                     "binding_resource": "%s %.1f %% of peak" % (top, util[top]), "utilisation_pct": util}
    json.dump(rec, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)


if __name__ == "__main__":
    {"report": report, "launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
