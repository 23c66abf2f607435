#!/usr/bin/env python
"""Turn raw Nsight Compute output (gpurun_out/, scratch) into the small tracked summaries under
profiles/ that DESIGN.md and bench.py cite.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  profiles/r01_launches.md
    python profiles/summarize.py kernel   gpurun_out/prof_x.ncu-rep   profiles/r01_align_fwd.md

`launches`: the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv` for one
bench.py run -> per-kernel launch count, total / mean duration and SHARE of the GPU time.
`kernel`: one `ncu --set full` report -> the metrics the roofline argument rests on (duration,
DRAM bytes read/written, DRAM and SM throughput, shared-memory wavefronts and bank conflicts,
occupancy limiters, issue-slot utilisation, the top stall reasons).  Needs `ncu` on PATH.
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, newline="")) if len(r) >= 15]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("rlod::", "").replace("void ", "")
        ns = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, {"n": 0, "ns": 0.0, "grid": r[gi], "block": r[bi]})
        a["n"] += 1
        a["ns"] += ns
    total = sum(a["ns"] for a in agg.values())
    out = ["| kernel | launches | total us | mean us | share | grid (last) | block |", "|---|---|---|---|---|---|---|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        out.append(f"| `{name}` | {a['n']} | {a['ns'] / 1e3:.1f} | {a['ns'] / 1e3 / a['n']:.1f} | "
                   f"{100 * a['ns'] / total:.1f}% | {a['grid']} | {a['block']} |")
    out.append(f"\ntotal GPU time over {sum(a['n'] for a in agg.values())} launches: {total / 1e3:.1f} us "
               "(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)")
    open(dst, "w").write(f"# launch list: {src}\n\n" + "\n".join(out) + "\n")
    print("\n".join(out))


KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__inst_executed_op_tma_st.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def kernel(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full summary: {src}\n"]
    js = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append(f"## `{d.get('Kernel Name', '?')}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        rec = {"kernel": d.get("Kernel Name")}
        for k in KEEP:
            if k in d:
                out.append(f"| {k} | {d[k]} | {u[k]} |")
                rec[k] = [d[k], u[k]]
        js.append(rec)
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    json.dump(js, open(dst.rsplit(".", 1)[0] + ".json", "w"), indent=1)
    print("\n".join(out))


def stalls(src, dst, top=25):
    """Per-SASS-instruction warp-stall samples of the FIRST kernel in the report (the source page)."""
    txt = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    name, hdr, data = rows[0][1], rows[1], [r for r in rows[2:starts[1]] if len(r) > 10]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {k: sum(int(r[ix[k]] or 0) for r in data) for k in keys}
    T = sum(tot.values()) or 1
    out = [f"# warp-stall sampling by SASS instruction: {src}\n", f"kernel `{name}`, {T} samples, {len(data)} SASS instructions\n",
           "stall reasons (% of samples): " + ", ".join(f"{k[6:]} {100 * v / T:.1f}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 200 > T),
           "", "| # | SASS | samples | % | dominant reason | executed |", "|---|---|---|---|---|---|"]
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = data[i]
        n = int(r[ix["# Samples"]] or 0)
        dom = max(keys, key=lambda k: int(r[ix[k]] or 0))
        out.append(f"| {i} | `{r[ix['Source']].strip()[:64]}` | {n} | {100 * n / T:.1f} | {dom[6:]} | {r[ix['Instructions Executed']]} |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "stalls": stalls}[sys.argv[1]](sys.argv[2], sys.argv[3])
