#!/usr/bin/env python
"""Turn an `ncu --set full` report of the GEMM kernels into the committed summaries:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_igemm_persist_summary.csv
writes the CSV (one column per captured launch) and profiles/roofline_traffic.json (mean DRAM bytes per launch)."""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(data))) + "\n")
        for h in KEEP:
            f.write(",".join([h, units[col[h]]] + [r[col[h]].replace(",", "") for r in data]) + "\n")
    tot = 0.0
    for h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += sum(float(r[col[h]].replace(",", "")) * SCALE[units[col[h]]] for r in data)
    dur = [float(r[col["gpu__time_duration.sum"]].replace(",", "")) for r in data]
    tens = [float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]].replace(",", "")) for r in data]
    j = {"dram_bytes_per_launch": round(tot / len(data)), "launches": len(data),
         "tensor_pipe_active_pct_time_weighted": round(sum(d * t for d, t in zip(dur, tens)) / sum(dur), 2),
         "source": f"{os.path.basename(out)} (ncu --set full --clock-control none of one train step: mean over the {len(data)} captured launches)"}
    json.dump(j, open(os.path.join(os.path.dirname(out), "roofline_traffic.json"), "w"), indent=1)
    print(j)


if __name__ == "__main__":
    main()
