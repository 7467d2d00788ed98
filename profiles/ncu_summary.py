"""Print the handful of ncu metrics we track for a kernel from a `--page raw --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70] if "Kernel Name" in hdr else "")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:75s} {r[i]:>16s} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            st.append((float(r[i] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
