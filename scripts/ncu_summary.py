"""Summarise `ncu --set full` reports (read here with `ncu -i <rep> --page raw --csv`): duration, grid, registers, DRAM bytes,
tensor-pipe utilisation (tcgen05 = utchmma path, mma.sync = hmma path), throughput percentages, top warp stall reasons.
    python scripts/ncu_summary.py gpurun_out/prof_*.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

PICK = [("duration us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"),
        ("dyn smem B", "launch__shared_mem_per_block_dynamic"), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("dram read MB", "dram__bytes_read.sum"), ("dram write MB", "dram__bytes_write.sum"),
        ("tensor utchmma bf16->fp32 %", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
        ("tensor hmma bf16->fp32 %", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
        ("tensor pipe cycles active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active")]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(rep, "(no kernels)")
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print(f"{rep}: {d.get('Kernel Name', '?')[:110]}")
        print("   " + " | ".join(f"{name} {d[key]}" for name, key in PICK if key in d and d[key] not in ("", "n/a")))
        stalls = sorted(((float(v), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k, v in d.items()
                         if "smsp__average_warps_issue_stalled_" in k and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")),
                        reverse=True)[:5]
        print("   stalls per issue: " + ", ".join(f"{n} {x:.2f}" for x, n in stalls))
