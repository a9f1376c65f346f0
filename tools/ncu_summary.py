#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (ncu --set full): duration, DRAM bytes, occupancy, issue rate.
Writes JSON (one object per captured launch) so the numbers can be committed under profiles/ and
bench.py can quote the draw kernel's measured DRAM traffic.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_mb",
    "dram__bytes_write.sum": "dram_write_mb",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_insts",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k, name in KEYS.items():
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                try:
                    v = float(v)
                except ValueError:
                    continue
                if name.endswith("_mb") and u.lower().startswith("kbyte"):
                    v /= 1e3
                if name.endswith("_mb") and u.lower() == "byte":
                    v /= 1e6
                if name == "time_us" and u == "ns":
                    v /= 1e3
                d[name] = round(v, 3)
        res.append(d)
    json.dump({"source": rep, "launches": res}, open(out, "w"), indent=1)
    for d in res:
        print(d)


if __name__ == "__main__":
    main()
