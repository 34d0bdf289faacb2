#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > /tmp/s.csv; python tools/ncu_lines.py /tmp/s.csv [top]
"""
import csv
import sys

rows = []
fname = None
cols = None
for r in csv.reader(open(sys.argv[1], newline="")):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].rsplit("/", 1)[-1]
    elif r[0] == "Line No":
        cols = {c: i for i, c in enumerate(r)}
    elif cols and r[0] not in ("", "Function Name", "File Name") and r[0].isdigit():
        try:
            rows.append((fname, int(r[0]), r[1].strip()[:110], int(r[cols["Instructions Executed"]]), int(r[cols["# Samples"]])))
        except (ValueError, IndexError):
            pass
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ti, ts = sum(r[3] for r in rows), sum(r[4] for r in rows)
print(f"total instr {ti}  samples {ts}")
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print(f"{100*r[3]/ti:5.1f}% i {100*r[4]/max(ts,1):5.1f}% s  {r[0]}:{r[1]:<4d} {r[2]}")
