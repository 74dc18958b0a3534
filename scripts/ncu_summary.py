#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the few numbers DESIGN.md / bench.py quote.

  python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--json]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and prints, per
captured launch: duration, DRAM bytes read/written, DRAM throughput, L2 hit rate, achieved
occupancy, registers, the issue-stall breakdown (top reasons) and instruction counts.
"""
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__bytes.sum.per_second", "dram_throughput"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_rate"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "registers_per_thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ_limit_regs_blocks"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem_blocks"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("smsp__cycles_active.avg", "smsp_cycles_active"),
    ("sm__cycles_elapsed.avg", "sm_cycles_elapsed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slot_util_pct"),
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12,
        "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "second": 1,
        "Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            d[h] = (v, u)
        res.append(d)
    return res


def num(d, k):
    if k not in d:
        return None
    v, u = d[k]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * UNIT.get(u, 1.0)


def summarise(d):
    s = {"kernel": d.get("Kernel Name", ("?", ""))[0]}
    for k, name in KEYS:
        s[name] = num(d, k)
    stalls = []
    for k in d:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            x = num(d, k)
            if x is not None:
                stalls.append((x, k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    stalls.sort(reverse=True)
    s["top_stalls_warps_per_issue"] = [(n, round(x, 3)) for x, n in stalls[:6]]
    if s["dram_read"] is not None and s["dram_write"] is not None:
        s["dram_bytes"] = s["dram_read"] + s["dram_write"]
    return s


def main():
    path = sys.argv[1]
    res = [summarise(d) for d in load(path)]
    if "--json" in sys.argv:
        print(json.dumps(res, indent=1))
        return
    for s in res:
        print(f"kernel: {s['kernel']}")
        for k, v in s.items():
            if k != "kernel":
                print(f"  {k:28s} {v}")


if __name__ == "__main__":
    main()
