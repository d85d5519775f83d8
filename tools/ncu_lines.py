"""Per-source-line instruction counts of one kernel from a .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_lines.py <rep> <kernel regex> [top N]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
ti = hdr.index("Thread Instructions Executed")
lines = []
fname = ""
for r in rows[:hdr_i]:
    if r and r[0] == "File Name": fname = r[1]
for r in rows[hdr_i + 1:]:
    if r and r[0] == "File Name": fname = r[1]; continue
    if r and r[0] not in ("", "Line No") and r[0].isdigit():
        try: lines.append((int(r[ci]), int(r[si]), int(r[ti]), fname.split("/")[-1], int(r[0]), r[1].strip()))
        except ValueError: pass
tot = sum(l[0] for l in lines); tots = sum(l[1] for l in lines)
print("total warp instructions %d, samples %d" % (tot, tots))
for l in sorted(lines, reverse=True)[:top]:
    print("%6.2f%% inst %6.2f%% samp  act %4.1f  %s:%d  %s" % (100.0 * l[0] / tot, 100.0 * l[1] / max(1, tots), l[2] / max(1, l[0]), l[3], l[4], l[5][:110]))
