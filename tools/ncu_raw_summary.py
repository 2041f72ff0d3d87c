"""Prints a compact per-kernel table of the metrics that matter from an `ncu --page raw --csv` dump.
    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_raw_summary.py raw.csv
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("==", r[ix["Kernel Name"]][:70], "id", r[ix["ID"]])
    for w in want:
        if w in ix:
            print(f"   {w:75s} {r[ix[w]]:>16s} {units[ix[w]]}")
