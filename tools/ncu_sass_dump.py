"""Per-SASS-instruction dump of an ncu report captured with --import-source on: address order, executed count, stall samples.
usage: ncu_sass_dump.py <report.ncu-rep> > out.txt"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if "Source" in r and any("Samples" in c for c in r))
h = rows[hdr]
col = {n: j for j, n in enumerate(h)}
samp = next(c for c in h if "Sampling (All" in c or c.startswith("# Samples"))
inst = next(c for c in h if c.startswith("Instructions Executed") and "Thread" not in c)
st = [c for c in h if c.startswith("stall_")]
for k, r in enumerate(rows[hdr + 1:]):
    if len(r) != len(h):
        continue
    s = int(float(r[col[samp]] or 0))
    top = sorted(((int(float(r[col[c]] or 0)), c[6:]) for c in st), reverse=True)[:4]
    print("%5d %10d %7d  %-70s %s" % (k, int(float(r[col[inst]] or 0)), s, r[col["Source"]].strip()[:70],
                                      " ".join("%s=%d" % (n, v) for v, n in top if v)))
