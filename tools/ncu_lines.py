"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "":  # a source-line summary row
        try:
            agg.append((cur_file, int(r[0]), r[1].strip(), int(r[4] or 0), int(r[7] or 0)))
        except ValueError:
            pass
tot_s = sum(a[3] for a in agg) or 1
tot_i = sum(a[4] for a in agg) or 1
print(f"total samples {tot_s}, total warp instructions {tot_i}")
print("---- by stall samples")
for f, ln, src, s, i in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{ln}  {src[:110]}")
print("---- by instructions")
for f, ln, src, s, i in sorted(agg, key=lambda a: -a[4])[:top]:
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{ln}  {src[:110]}")
