"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small text file for profiles/."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print(f"== kernel: {r[name_i]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>20s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  issue-stall breakdown (warps stalled per issue-active cycle): " +
              ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
