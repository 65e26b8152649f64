"""Condenses an .ncu-rep into the handful of numbers we track (run here, no GPU needed):
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--out profiles/r01_x.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:70], d.get("Grid Size", "?"), d.get("Block Size", "?")))
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append("  %-70s %-16s %s" % (k, units[hdr.index(k)], d[k]))
        st = [(h, float(d[h])) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and d.get(h)]
        st.sort(key=lambda x: -x[1])
        lines.append("  stalls (warps per issue): " + ", ".join("%s=%.2f" % (h.split("issue_stalled_")[1].split("_per_")[0], v) for h, v in st[:8]))
    s = "\n".join(lines)
    print(s)
    if out:
        with open(out, "w") as fh:
            fh.write("source: %s (ncu --set full --clock-control none)\n" % rep + s + "\n")


if __name__ == "__main__":
    main()
