#!/usr/bin/env python
"""Summarise `ncu --page source --csv` of one kernel: top SASS lines by stall samples, shared
wavefronts and global sectors.  usage: ncu_source_summary.py report.ncu-rep kernel_regex [instance]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
import re
blocks = [b for b in out.split('"Kernel Name",')[1:] if re.search(rx, b.split("\n")[0])]
blk = blocks[inst]
lines = blk.split("\n")
print("kernel:", lines[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot_samples = sum(int(r[col["# Samples"]]) for r in data)
tot_wf = sum(int(r[col["L1 Wavefronts Shared"]]) for r in data)
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in data)
print(f"total samples {tot_samples}, shared wavefronts {tot_wf}, warp instructions {tot_inst}, SASS lines {len(data)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[col[s]]) for r in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
print("\n-- top lines by samples")
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:28]:
    top = sorted(((int(r[col[s]]), s) for s in stalls), reverse=True)[:2]
    print(f"{int(r[col['# Samples']]):7d} inst={int(r[col['Instructions Executed']]):9d} wf={int(r[col['L1 Wavefronts Shared']]):9d}/{int(r[col['L1 Wavefronts Shared Ideal']]):9d} "
          f"sect={int(r[col['L2 Theoretical Sectors Global']]):9d} {r[col['Source']].strip()[:70]:70s} {top}")
print("\n-- shared-memory lines")
for r in sorted(data, key=lambda r: -int(r[col["L1 Wavefronts Shared"]]))[:14]:
    if int(r[col["L1 Wavefronts Shared"]]) == 0:
        break
    print(f"wf={int(r[col['L1 Wavefronts Shared']]):9d} ideal={int(r[col['L1 Wavefronts Shared Ideal']]):9d} inst={int(r[col['Instructions Executed']]):9d} {r[col['Source']].strip()[:80]}")
