"""Print the warp-stall breakdown and pipe / memory utilisation of every launch in an `ncu --page raw --csv` export."""
import csv
import sys
r = list(csv.reader(open(sys.argv[1])))
hdr, units, rows = r[0], r[1], r[2:]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
keys = [i for i, h in enumerate(hdr) if ("issue_stalled" in h and h.endswith("_per_warp_active.pct")) or h in (
    "gpu__time_duration.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active")]
name = hdr.index("Kernel Name")
for row in rows:
    if pat and pat not in row[name]:
        continue
    print("==", row[name][:90])
    vals = []
    for i in keys:
        try:
            v = float(row[i].replace(",", ""))
        except ValueError:
            continue
        vals.append((hdr[i], v, units[i]))
    for h, v, u in vals:
        if "issue_stalled" in h and v < 2.0:
            continue
        print(f"   {h:95s} {v:14.2f} {u}")
