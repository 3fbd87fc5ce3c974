#!/usr/bin/env python
"""Per-shape DRAM traffic of the GEMM kernel from an `ncu --set full` capture -> profiles/ncu_gemm_traffic.json.

  (GPU box)  ncu --set full --clock-control none -k regex:gemm_f16 -c 400 -o gpurun_out/prof_gemm \\
                 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-roofline \\
                 --no-e2e --gemm-log gpurun_out/gemm_order.json
  (here)     python tools/ncu_traffic.py gpurun_out/prof_gemm.ncu-rep gpurun_out/gemm_order.json [more pairs ...]

`bench.py --gemm-log` records the shape key of every GEMM launch of the process in launch order, so the n-th captured
`gemm_f16*` kernel IS the n-th entry (no skip, the capture starts at the first launch).  For every shape: median duration,
DRAM bytes read + written (dram__bytes_read.sum + dram__bytes_write.sum), tensor-pipe activity.  bench.py fills
`roofline.traffic` of its dominant shape from this file.
"""
import csv
import json
import statistics
import subprocess
import sys
from pathlib import Path

M = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        rec = {"name": r[col["Kernel Name"]]}
        for k, m in M.items():
            rec[k] = float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
        yield rec


def main(argv):
    shapes, sources = {}, []
    for rep, order in zip(argv[0::2], argv[1::2]):
        keys = json.loads(Path(order).read_text())
        recs = [r for r in rows_of(rep) if "gemm_f16" in r["name"]]
        sources.append(f"{Path(rep).name}: {len(recs)} gemm launches captured of {len(keys)} logged")
        for key, r in zip(keys, recs):
            shapes.setdefault(key, []).append(r)
    out = {"source": "ncu --set full --clock-control none -k regex:gemm_f16 of bench.py --steps 1 --warmup 1 (per-launch, "
                     "cold-cache, serialised); " + "; ".join(sources),
           "shapes": {k: {"launches": len(v), "dram_bytes": statistics.median(x["rd"] + x["wr"] for x in v),
                          "dram_read_bytes": statistics.median(x["rd"] for x in v),
                          "dram_write_bytes": statistics.median(x["wr"] for x in v),
                          "dur_us": statistics.median(x["dur"] for x in v),
                          "tensor_pipe_pct": statistics.median(x["tensor"] for x in v)} for k, v in shapes.items()}}
    dst = Path(__file__).resolve().parent.parent / "profiles" / "ncu_gemm_traffic.json"
    dst.write_text(json.dumps(out, indent=1, sort_keys=True))
    for k, v in sorted(out["shapes"].items(), key=lambda kv: -kv[1]["dur_us"] * kv[1]["launches"])[:16]:
        print(f"{k}: {v['launches']} x {v['dur_us']:.1f} us, DRAM {v['dram_bytes'] / 1e6:.1f} MB, tensor {v['tensor_pipe_pct']:.1f} %")
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1:])
