"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + top stall reasons."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__sass_inst_executed_op_shared_ld.sum"]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        if len(vals) != len(hdr):
            continue
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d: print(f"  {k:70s} {d[k]:>18s} {u[k]}")
        stalls = [(k, float(d[k].replace(',', ''))) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k]]
        if not stalls:
            stalls = [(k, float(d[k].replace(',', ''))) for k in hdr if "warp_issue_stalled" in k and k.endswith(".ratio") and d[k]]
        for k, v in sorted(stalls, key=lambda x: -x[1])[:8]:
            print(f"  stall {k:80s} {v:10.3f}")
        print()
if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p); print()
