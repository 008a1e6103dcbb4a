"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_rd_sectors"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
    ("smsp__inst_executed.sum", "inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print(f"# {rep}: one block per profiled launch (ncu --set full --clock-control none)")
    for r in body:
        name = r[ki]
        print(name[:150])
        for m, short in WANT:
            if m in hdr:
                i = hdr.index(m)
                print(f"    {short:22s} {r[i]:>18s} {units[i]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
