"""Key metrics of an .ncu-rep of the d > 4 kernels (tuning aid / profiles/): python scripts/ncu_mid_summary.py rep [steps]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
W = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
     ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
     ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
     ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_inst_%"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
     ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
     ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"),
     ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall_short_sb"),
     ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall_math_throttle"),
     ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall_mio_throttle"),
     ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall_wait"),
     ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall_barrier"),
     ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall_lg_throttle"),
     ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "stall_no_inst"),
     ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall_not_selected"),
     ("smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "stall_dispatch")]
print(f"# {rep} (ncu --set full --clock-control none); per-step figures for {steps:.0f} time steps per launch")
for r in rows[2:]:
    print(r[idx["Kernel Name"]][:120])
    for m, s in W:
        if m in idx:
            v = r[idx[m]]
            extra = ""
            if s in ("warp_inst", "smem_wavefronts"):
                extra = f"   ({float(v) / steps:.0f} per time step)"
            print(f"    {s:22s} {v:>18s} {units[idx[m]]}{extra}")
