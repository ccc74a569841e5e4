"""Summarise an .ncu-rep (read here, on the CPU box) into a small markdown file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r1_x.md "title / what was run" """
import csv, subprocess, sys, io, collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second"]

def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + title, "", "source: `%s` (ncu --set full --clock-control none --import-source on)" % rep, ""]
    for data in rows[2:]:
        name = data[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines += ["## " + name, "", "| metric | value | unit |", "|---|---|---|"]
        for i, h in enumerate(hdr):
            if h in KEYS or any(h.endswith(k) for k in ("tensor_cycles_active.avg.pct_of_peak_sustained_active",)):
                lines.append("| %s | %s | %s |" % (h, data[i], units[i]))
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    byop, tot = collections.Counter(), 0
    for r in srows[2:]:
        try:
            n = int(r[5])
        except Exception:
            continue
        toks = r[1].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += n
        tot += n
    if tot:
        lines += ["## instruction mix of the first kernel (warp-level, from the source page)", "", "| opcode | share |", "|---|---|"]
        lines += ["| %s | %.1f %% |" % (k, 100.0 * v / tot) for k, v in byop.most_common(14)]
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)

if __name__ == "__main__":
    main()
