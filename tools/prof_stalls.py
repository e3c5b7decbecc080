"""Per source line: samples by stall reason, for one kernel section of an ncu `--page source --csv --print-source cuda,sass` dump.
   python tools/prof_stalls.py dump.csv FUNCTION_SUBSTRING [top] [reason]"""
import csv, sys
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
reason = sys.argv[4] if len(sys.argv) > 4 else "stall_long_sb"
sec = ""; hdr = None; cur = None; data = {}
for r in rows:
    if len(r) > 5 and r[0] == "Line No":
        # a new kernel section starts when the first source file header repeats
        hdr = r
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        sec = r[1]
        continue
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if hdr is None or want not in sec or len(r) < len(hdr) - 2 or r[2] != "-":
        continue
    try:
        key = (cur, int(r[0]), r[1].strip()[:80])
        data[key] = (int(float(r[hdr.index("# Samples")])), int(float(r[hdr.index(reason)] or 0)), int(float(r[hdr.index("Instructions Executed")])))
    except ValueError:
        pass
tot = sum(v[0] for v in data.values()) or 1
totr = sum(v[1] for v in data.values()) or 1
print("kernel %s: samples %d, %s %d (%.1f%%)" % (want, tot, reason, totr, 100.0 * totr / tot))
for k, v in sorted(data.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% of %s  (%4.1f%% of all samples)  %s:%d  %s" % (100.0 * v[1] / totr, reason, 100.0 * v[0] / tot, k[0], k[1], k[2]))
