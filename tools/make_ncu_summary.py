"""Round summaries from the `ncu --page raw --csv` exports of tools/run_call.sh:
    python tools/make_ncu_summary.py r2    ->  profiles/r2_ncu_summary.csv, profiles/r2_ncu_traffic.json
"""
import csv
import json
import os
import re
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp_insts"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def short(name):
    m = re.match(r"(void )?(gww::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(3) + (m.group(4) or "")) if m else name[:60]


out_rows = []
traffic = {"source": f"profiles/{tag}_ncu_summary.csv (ncu --set full --clock-control none; whisper-base, 256 det-windows per launch; "
                     "tools/run_call.sh, tools/make_ncu_summary.py)"}
GEMM_ROLE = ["gemm_qkv", "gemm_out_proj", "gemm_fc1", "gemm_fc2"]   # capture order: layer 1 of tools/profile_step.py
for cap in ("attn", "logmel", "gemm", "qfront"):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_{cap}_raw.csv")
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        rec = {"capture": cap, "kernel": short(r[ix["Kernel Name"]])}
        for metric, key in WANT:
            if metric in ix:
                v = float(r[ix[metric]].replace(",", "") or 0)
                u = units[ix[metric]]
                if key in ("dram_read", "dram_write"):
                    v *= UNIT.get(u, 1.0)
                elif key == "time":
                    v *= UNIT.get(u, 1.0)
                rec[key] = v
        rec["dram_bytes"] = rec.get("dram_read", 0) + rec.get("dram_write", 0)
        if cap == "gemm" and n < 4:
            rec["role"] = GEMM_ROLE[n]
            traffic[GEMM_ROLE[n]] = {"dram_bytes_per_launch": rec["dram_bytes"], "det_windows": 256}
        if cap == "attn":
            rec["role"] = "attention"
            traffic["attention_persist_kernel"] = {"dram_bytes_per_launch": rec["dram_bytes"], "det_windows": 256}
        if cap == "qfront":
            k = rec["kernel"]
            role = ("qscan_tiles" if k.startswith("qscan_tiles") else "qscan_interp" if k.startswith("qscan_interp") else
                    "qadapter_conv1" if k.startswith("qadapter_conv1") else "qadapter_pool" if k.startswith("qadapter_pool") else
                    "qadapter_conv2" if "<16, 32" in k else "qadapter_conv3" if "<32, 64" in k else k)
            rec["role"] = role
            traffic[role] = {"dram_bytes_per_launch": rec["dram_bytes"], "det_windows": 256}
        if cap == "logmel":
            rec["role"] = "logmel"
            traffic["logmel_kernel"] = {"dram_bytes_per_launch": rec["dram_bytes"], "det_windows": 256}
        out_rows.append(rec)
keys = ["capture", "role", "kernel", "time", "sm_clock", "tensor_pipe_pct", "xu_pipe_pct", "fma_pipe_pct", "alu_pipe_pct",
        "fp64_pipe_pct", "issue_active_pct", "dram_read", "dram_write", "dram_bytes", "dram_pct", "l1tex_pct", "l2_pct", "regs",
        "grid", "warp_insts"]
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on, one launch each (time in ms, bytes in B, clock in GHz);\n")
    f.write("# commands: tools/profile_step.py --windows 128 (whisper-base, 256 det-windows) and tools/profile_mlgwsc.py (256 windows)\n")
    w = csv.writer(f)
    w.writerow(keys)
    for rec in out_rows:
        w.writerow([("%.6g" % rec[k]) if isinstance(rec.get(k), float) else rec.get(k, "") for k in keys])
json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w"), indent=1)
for rec in out_rows:
    print(rec["capture"], rec.get("role", ""), rec["kernel"][:50], "ms=%.4f tensor=%.1f xu=%.1f fp64=%.1f issue=%.1f dram=%.3g GB (%.1f%%)" % (
        rec["time"], rec.get("tensor_pipe_pct", 0), rec.get("xu_pipe_pct", 0), rec.get("fp64_pipe_pct", 0), rec.get("issue_active_pct", 0),
        rec["dram_bytes"] / 1e9, rec.get("dram_pct", 0)))
