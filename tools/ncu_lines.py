"""Aggregate an ncu source page (cuda,sass CSV) per CUDA source line: executed instructions and stall samples.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv; python tools/ncu_lines.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 8 and r[0].isdigit() and r[2] == "-":
        try:
            data.append((cur_file, int(r[0]), r[1].strip(), int(r[7]), int(r[4])))
        except ValueError:
            pass
tot = sum(d[3] for d in data)
tots = sum(d[4] for d in data)
print("total executed warp-instructions %d, stall samples %d" % (tot, tots))
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%-14s %4d  exec %5.1f%%  stall %5.1f%%  %s" % (d[0], d[1], d[3] / tot * 100, d[4] / tots * 100, d[2][:110]))
