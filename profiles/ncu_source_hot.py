#!/usr/bin/env python3
"""Per-CUDA-source-line stall samples from an ncu report captured with --import-source on (-lineinfo build):
  python profiles/ncu_source_hot.py <report.ncu-rep> [top N]
Aggregates `ncu --page source --print-source cuda,sass --csv` over files and lines."""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, acc, total, kern = None, None, {}, 0, None
for r in rows:
  if not r:
    continue
  if r[0] == "File Path":
    cur_file = r[1].split("/")[-1]; continue
  if r[0] == "Function Name":
    kern = r[1][:100]; continue
  if r[0] == "Line No":
    hdr = {h: i for i, h in enumerate(r)}; continue
  if hdr is None or not r[0].isdigit() or r[hdr["Address"]] != "-":
    continue   # SASS rows under a source line carry an address
  try:
    smp = int(r[hdr["# Samples"]]); ins = int(r[hdr["Instructions Executed"]])
  except (ValueError, KeyError):
    continue
  stalls = {k[6:]: int(r[i]) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k and r[i].isdigit() and int(r[i]) > 0}
  key = (cur_file, int(r[0]))
  a = acc.setdefault(key, [0, 0, r[1].strip()[:110], {}])
  a[0] += smp; a[1] += ins
  for k, v in stalls.items():
    a[3][k] = a[3].get(k, 0) + v
  total += smp
print(f"# {kern}\n# total samples {total}")
for (f, ln), (smp, ins, src, st) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
  top3 = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
  print(f"{100.0 * smp / max(total, 1):5.1f}%  inst {ins:>9d}  {f}:{ln:<4d} {src}\n         [{top3}]")
