"""Key raw-page metrics of an .ncu-rep as text: python tools/ncu_summary.py report.ncu-rep > profiles/<name>_ncu_summary.txt"""
import csv, subprocess, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
ix = {n: i for i, n in enumerate(h)}
print("# kernel:", v[ix["Kernel Name"]] if "Kernel Name" in ix else "?")
for n in KEEP:
    if n in ix:
        print(f"{n:90s} {v[ix[n]]:>20s} {u[ix[n]]}")
for i, n in enumerate(h):
    if "issue_stalled" in n and "per_issue_active" in n and float(v[i] or 0) > 0.02:
        print(f"{n:90s} {v[i]:>20s} {u[i]}")
for extra in sys.argv[2:]:
    print("#", extra)
