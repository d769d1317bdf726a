"""Warp-stall samples and executed instructions by opcode from an ncu report captured with --import-source on.
usage: ncu_stalls.py <report.ncu-rep> [top_n]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for i, r in enumerate(rows):
        if "Source" in r and any("Samples" in c for c in r):
            hdr = i
            break
    h = rows[hdr]
    col = {name: j for j, name in enumerate(h)}
    samp = next(c for c in h if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
    inst = next(c for c in h if c.startswith("Instructions Executed") and "Thread" not in c)
    stall_cols = [c for c in h if c.startswith("stall_")]
    agg = collections.defaultdict(lambda: collections.Counter())
    tot_s = tot_i = 0
    for r in rows[hdr + 1:]:
        if len(r) != len(h):
            continue
        src = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip())
        op = src.split()[0].split(".")[0] if src else "?"
        s = int(float(r[col[samp]] or 0)); n = int(float(r[col[inst]] or 0))
        agg[op]["samples"] += s; agg[op]["inst"] += n
        tot_s += s; tot_i += n
        for c in stall_cols:
            agg[op][c] += int(float(r[col[c]] or 0))
    want = ["stall_short_sb", "stall_wait", "stall_long_sb", "stall_selected", "stall_math", "stall_dispatch", "stall_not_selected",
            "stall_branch_resolving", "stall_barrier", "stall_lg", "stall_mio"]
    want = [w for w in want if w in stall_cols]
    print("total samples %d, warp-instructions executed %d" % (tot_s, tot_i))
    print("%-10s %9s %6s %12s " % ("op", "samples", "%", "inst_exec") + " ".join("%9s" % w[6:15] for w in want))
    for op, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        print("%-10s %9d %5.1f%% %12d " % (op, c["samples"], 100.0 * c["samples"] / max(tot_s, 1), c["inst"]) +
              " ".join("%9d" % c[w] for w in want))


if __name__ == "__main__":
    main()
