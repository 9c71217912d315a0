#!/usr/bin/env python
"""Reads `ncu --page raw --csv` exports under profiles/ and writes profiles/traffic.json (DRAM bytes per launch, read by
bench.py for roofline.traffic) plus a compact per-kernel summary (profiles/summary.json).  The raw page scales units per
metric (byte, Kbyte, Mbyte, Gbyte; ns, us, ms): the units row is honoured."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6,
         "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
KEEP = {
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.per_cycle_active": "issue_per_cycle_per_smsp",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_popc_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__grid_size": "grid", "launch__registers_per_thread": "registers", "launch__waves_per_multiprocessor": "waves",
}


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            out[h] = float(v.replace(",", "")) * SCALE.get(u, 1.0)
        except ValueError:
            out[h] = v
    return out


def main():
    # (file, traffic key, pairs per launch of the captured command)
    captures = [a.split(":") for a in sys.argv[1:]]
    traffic, summary = {}, {}
    tpath = os.path.join(PROF, "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for fname, key, pairs in captures:
        d = read(os.path.join(PROF, fname))
        dram = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        traffic[key] = {"dram_bytes": dram, "pairs_per_launch": int(pairs),
                        "source": f"profiles/{fname} (ncu --set full, 1 launch)"}
        s = {"kernel": str(d.get("Kernel Name", ""))[:120], "duration_us": d["gpu__time_duration.sum"] * 1e6,
             "dram_read_MB": d["dram__bytes_read.sum"] / 1e6, "dram_write_MB": d["dram__bytes_write.sum"] / 1e6,
             "pairs_per_launch": int(pairs), "source": fname}
        for k, name in KEEP.items():
            if k in d:
                s[name] = d[k]
        summary[key] = s
    json.dump(traffic, open(tpath, "w"), indent=1)
    spath = os.path.join(PROF, "summary.json")
    old = json.load(open(spath)) if os.path.exists(spath) else {}
    old.update(summary)
    json.dump(old, open(spath, "w"), indent=1)
    for k, s in summary.items():
        print(k, json.dumps(s))


if __name__ == "__main__":
    main()
