"""ncu report -> the text summary committed under profiles/ (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/r2_tcf_fix.ncu-rep > profiles/r2_tcf_full_summary.txt
Prints, per profiled launch, a fixed list of raw metrics (time, DRAM / L2 / shared traffic, pipe utilisation, issue
and stall breakdown, launch geometry)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s (ncu --set full --clock-control none; raw page, selected metrics; times are under the profiler)" % rep)
    for r in rows[2:]:
        print("kernel: %s   grid %s block %s" % (r[col["Kernel Name"]], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        for k in KEYS:
            hits = [h for h in hdr if h == k or h.endswith("." + k)]
            for h in hits[:1]:
                print("  %-95s %16s %s" % (k, r[col[h]], units[col[h]]))


if __name__ == "__main__":
    main()
