"""Summarise `ncu --set full` reports into the text files committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_X.ncu-rep [...]  -> profiles/r2_<name>_summary.txt (ROUND=rN)

Reads the raw page of each report (`ncu -i ... --page raw --csv`) and keeps the metrics the
roofline discussion in DESIGN.md uses: duration, DRAM bytes, issue activity, pipe
utilisation, occupancy, warp stall reasons."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def summarize(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    d = {n: (v, u) for n, u, v in zip(names, units, vals)}
    out = ["report: %s" % os.path.basename(rep), "kernel: %s" % d.get("Kernel Name", ("?", ""))[0],
           "grid %s block %s" % (d.get("Grid Size", ("?", ""))[0], d.get("Block Size", ("?", ""))[0]), ""]
    for k in KEEP:
        if k in d:
            out.append("%-70s %18s %s" % (k, d[k][0], d[k][1]))
    out.append("")
    out.append("warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active, ratio):")
    stalls = sorted(((float(v[0].replace(",", "")), k) for k, v in d.items()
                     if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[0] not in ("", "n/a")),
                    reverse=True)
    for v, k in stalls[:8]:
        out.append("  %-40s %6.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    try:
        rd = float(d["dram__bytes_read.sum"][0].replace(",", "")); wr = float(d["dram__bytes_write.sum"][0].replace(",", ""))
        ur, uw = d["dram__bytes_read.sum"][1], d["dram__bytes_write.sum"][1]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = rd * scale.get(ur, 1) + wr * scale.get(uw, 1)
        ns = float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(d["gpu__time_duration.sum"][1], 1)
        out.append("")
        out.append("traffic: dram read + write = %.0f bytes per launch; %.1f GB/s over the (profiled, unlocked-clock) duration" % (tot, tot / ns))
    except Exception as e:
        out.append("traffic: n/a (%s)" % e)
        tot = None
    return "\n".join(out) + "\n", tot


if __name__ == "__main__":
    for rep in sys.argv[1:]:
        name = os.path.basename(rep).replace(".ncu-rep", "").replace("prof_", "")
        text, tot = summarize(rep)
        out_dir = os.environ.get("SUMMARY_DIR", os.path.join(ROOT, "profiles"))      # on the GPU box: gpurun_out/summaries (reports are too big to bring back)
        os.makedirs(out_dir, exist_ok=True)
        dst = os.path.join(out_dir, "%s_%s_summary.txt" % (os.environ.get("ROUND", "r2"), name))
        open(dst, "w").write(text)
        print(dst, tot)
