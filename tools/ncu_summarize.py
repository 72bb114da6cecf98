#!/usr/bin/env python
"""Turns the ncu outputs of tools/gpu_round.sh (gpurun_out/) into the small tracked summaries under profiles/:
   launches_<tag>.csv          -> profiles/<tag>_launches.csv (copy) + <tag>_launches_summary.csv (time share per kernel)
   prof_<tag>.ncu-rep (full)   -> profiles/<tag>_ncu_full_summary.json (per kernel: duration, DRAM bytes, pipe/stall
                                  metrics) and profiles/traffic.json (DRAM bytes per particle, read by bench.py)
usage: python tools/ncu_summarize.py <tag> <particles>"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, npart = sys.argv[1], float(sys.argv[2])
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list
src = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f"{tag}_launches.csv"))
    lines = [ln for ln in open(src) if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launches_summary.csv"), "w") as f:
        f.write(f"# ncu launch list of bench.py --steps 2 --warmup 3 --no-e2e --no-cpu ({npart:.3g} particles), launches 125..185\n")
        f.write("# gpu__time_duration.sum; per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write("launches,total_ms,share,kernel\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n},{ms:.3f},{100 * ms / tot:.1f}%,{k}\n")
    print(open(os.path.join(P, f"{tag}_launches_summary.csv")).read())

# ---- full capture
rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "launch__grid_size",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
    units = dict(zip(hdr, rows[1]))
    out, traffic = [], {}
    grp = {"advect_lean": "move", "advect_full": "move", "scatter": "sort", "cell_collide": "collide"}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = OrderedDict(kernel=d["Kernel Name"].split("(")[0])
        for k in keep:
            if k in d:
                e[k] = f"{d[k]} {units.get(k, '')}".strip()
        try:
            rd, wr = float(d["dram__bytes_read.sum"].replace(",", "")), float(d["dram__bytes_write.sum"].replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units["dram__bytes_read.sum"]]
            e["dram_bytes_per_particle"] = round((rd + wr) * scale / npart, 2)
            for k, g in grp.items():
                if k in e["kernel"]:
                    traffic[g] = traffic.get(g, 0.0) + (rd + wr) * scale / npart
        except Exception:
            pass
        out.append(e)
    json.dump(out, open(os.path.join(P, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    json.dump({"source": f"ncu --set full --clock-control none, profiles/{tag}_ncu_full_summary.json, C5 strip {npart:.3g} particles",
               "dram_bytes_per_particle": traffic}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
