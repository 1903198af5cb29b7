#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep captures into the small text summaries committed under profiles/.
usage: python profiles/summarize.py <report.ncu-rep> <out.txt> [kernel-substring]"""
import csv
import collections
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_atom.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    rep, out = sys.argv[1], sys.argv[2]
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none, report {rep.split('/')[-1]} (metrics per launch)"]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if filt and filt not in name:
            continue
        lines.append(f"\n== {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:70s} {r[i]:>18s} {units[i]}")
        st = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                st.append((num(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        tot = sum(v for v, _ in st) or 1
        lines.append("warp stall samples: " + ", ".join(f"{n} {100*v/tot:.1f}%" for v, n in sorted(st, reverse=True)[:8]))
    # hottest source lines (needs -lineinfo + --import-source on)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur_file = cur_fn = hdr2 = None
    agg = collections.OrderedDict()
    for r in csv.reader(src.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if len(r) == 2 and r[0] == "Function Name":
            cur_fn = r[1]; continue
        if r and r[0] == "Line No":
            hdr2 = r; continue
        if hdr2 and len(r) >= 8 and r[0] != "":
            if filt and filt not in (cur_fn or ""):
                continue
            key = ((cur_fn or "")[:40], cur_file, r[0], r[1][:96])
            a = agg.setdefault(key, [0, 0, 0])
            a[0] += num(r[hdr2.index("Instructions Executed")]); a[1] += num(r[hdr2.index("# Samples")])
            a[2] += num(r[hdr2.index("Thread Instructions Executed")])
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    lines.append("\n# hottest source lines (all captured launches): warp instructions, share, stall-sample share, active threads/instruction")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        lines.append(f"{k[0][:28]:28s} {k[1][:18]:18s} L{k[2]:>4s} {a[0]/1e6:9.2f}M {100*a[0]/tot:5.1f}% samp {100*a[1]/tots:5.1f}% thr {a[2]/max(a[0],1):5.1f} | {k[3]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
