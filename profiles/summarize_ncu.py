#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep, from `ncu --set full --import-source on`) into the text/JSON
files committed under profiles/.  Usage: python profiles/summarize_ncu.py <report.ncu-rep> <out-prefix>"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def source(rep, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    cur, agg, tot_i, tot_s, hot, cold = None, {}, 0, 0, 0, 0
    line = None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] in ("Function Name", "Line No"):
            continue
        elif r[0].isdigit():
            try:
                inst, samp = int(r[7]), int(r[4])
            except ValueError:
                continue
            line = (cur, int(r[0]), r[1].strip()[:96])
            a = agg.setdefault(line, [0, 0])
            a[0] += inst
            a[1] += samp
            tot_i += inst
            tot_s += samp
        elif r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
            try:
                if int(r[7]) > 0:
                    hot += 1
                else:
                    cold += 1
            except ValueError:
                pass
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]
    return tot_i, tot_s, hot, cold, rows


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]] for r in data]
    summary = {"report": rep.split("/")[-1], "kernels": names, "launches_profiled": len(data), "metrics": {}}
    with open(prefix + ".txt", "w") as f:
        f.write(f"ncu summary of {rep.split('/')[-1]} ({len(data)} launch(es) of {names[0]})\n")
        f.write("captured with: ncu --set full --clock-control none --import-source on (see DESIGN.md, Measurement)\n\n")
        for k in KEYS + sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
            if k in idx:
                vals = [r[idx[k]] for r in data]
                f.write(f"{k:88s} {units[idx[k]]:16s} {vals}\n")
                summary["metrics"][k] = {"unit": units[idx[k]], "values": vals}
        tot_i, tot_s, hot, cold, rows = source(rep)
        f.write(f"\nsource page: {tot_i} warp instructions executed, {tot_s} stall samples; "
                f"{hot} distinct SASS instructions executed at least once, {cold} never executed\n")
        f.write("top source lines by executed warp instructions (inst, % of total, stall samples):\n")
        for (fn, ln, txt), (inst, samp) in rows:
            f.write(f"  {fn}:{ln:<4d} {inst:10d} {100 * inst / max(tot_i, 1):5.1f}% {samp:6d}  | {txt}\n")
        summary["source"] = {"warp_instructions": tot_i, "stall_samples": tot_s, "hot_sass": hot, "cold_sass": cold}
    with open(prefix + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    print("wrote", prefix + ".txt")


if __name__ == "__main__":
    main()
