#!/usr/bin/env python
"""Condense an `ncu --set full` report (.ncu-rep) into a small CSV for profiles/.

usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/rNN_x_summary.csv

One row per captured launch: kernel, grid, block, registers, duration, DRAM bytes read/written,
DRAM / L2 / SM throughput as % of peak, FP64 pipe activity, achieved occupancy.
"""
import csv
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_B"),
    ("gpu__time_duration.sum", "duration_us"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu_dram_pct"),
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram_active_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__inst_executed.sum", "inst"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    sel = [(head.index(k), name, k) for k, name in COLS if k in head]
    # every FP64-tensor (DMMA) counter the capture holds, whatever this ncu version calls them
    known = {k for k, _ in COLS}
    sel += [(i, k, k) for i, k in enumerate(head) if "dmma" in k.lower() and k not in known]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([name + ("[%s]" % units[i] if units[i] else "") for i, name, _ in sel])
        for r in data:
            vals = []
            for i, name, _ in sel:
                v = r[i]
                if name == "kernel":
                    v = v.split("::")[-1].split("(")[0]
                vals.append(v)
            w.writerow(vals)
    print("wrote", out, len(data), "launches")


if __name__ == "__main__":
    main()
