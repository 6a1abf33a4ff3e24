#!/usr/bin/env python
"""One line per captured kernel from `ncu -i X.ncu-rep --page raw --csv`: duration, DRAM bytes, instructions,
issue utilisation, lanes per instruction, occupancy, the largest stall reason.  usage: kernel_table.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1, "us": 1, "nsecond": 1e-3, "ns": 1e-3, "msecond": 1e3, "ms": 1e3}


def val(r, name):
    return float(r[idx[name]].replace(",", "")) * MUL.get(units[idx[name]], 1)


stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("| kernel | us | DRAM MB (R+W) | warp instr | issue active | lanes / instr | warps active | top stall (per issue) | regs | grid x block |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("bendy::", "")
    top = max(stalls, key=lambda h: float(r[idx[h]].replace(",", "") or 0))
    tname = top[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
    print(f"| `{name}` | {val(r, 'gpu__time_duration.sum'):.1f} | {(val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')) / 1e6:.1f} "
          f"| {val(r, 'smsp__inst_executed.sum') / 1e6:.2f} M | {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} % "
          f"| {val(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} % "
          f"| {tname} {float(r[idx[top]].replace(',', '')):.1f} | {int(val(r, 'launch__registers_per_thread'))} "
          f"| {int(val(r, 'launch__grid_size'))} x {int(val(r, 'launch__block_size'))} |")
