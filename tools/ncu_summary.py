#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into one line per captured launch: duration, DRAM traffic, tensor-pipe %,
issue/ALU/XU utilisation, achieved occupancy.  Runs where there is no GPU:  python tools/ncu_summary.py rep.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("dur_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_inst_pct", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("fma_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("alu_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_pct", "sm__issue_active.avg.pct_of_peak_sustained_active"),
    ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("l2_pct", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
]


def main(path, grep=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    if grep:
        for h in hdr:
            if grep in h:
                print(h, units[col[h]], [r[col[h]] for r in rows[2:]][:6])
        return
    for r in rows[2:]:
        name = r[col["Kernel Name"]][:70]
        parts = []
        for short, m in WANT:
            if m in col:
                v = r[col[m]]
                u = units[col[m]]
                try:
                    f = float(v.replace(",", ""))
                    if m.startswith("dram__bytes"):
                        f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0) * f
                    if m.startswith("gpu__time"):
                        f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0) * f
                    v = f"{f:.1f}"
                except ValueError:
                    pass
                parts.append(f"{short}={v}")
        print(name, " ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
