"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv, sys
csv.field_size_limit(10**9)
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples"); i_n = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[2] != "-":      # SASS row
        continue
    try:
        data.append((int(float(r[i_s])), int(float(r[i_n])), cur_file, r[0], r[1].strip()[:100]))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
print("total samples %d, warp instructions %d" % (tot, toti))
for d in sorted(data, reverse=True)[:top]:
    print("%6.2f%% samp %6.2f%% inst  %s:%s  %s" % (100 * d[0] / tot, 100 * d[1] / toti, d[2], d[3], d[4]))
