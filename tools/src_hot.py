#!/usr/bin/env python
"""Hot CUDA source lines from `ncu -i rep --page source --print-source cuda,sass --csv --kernel-name regex:X > f.csv`."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname, hdr, col, out = "", None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            col = {}
            for i, h in enumerate(hdr):
                col.setdefault(h, i)
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":      # keep the per-line aggregate rows (Address == '-')
            continue
        try:
            s = int(r[col["# Samples"]] or 0)
            ex = int(r[col["Instructions Executed"]] or 0)
        except ValueError:
            continue
        reasons = sorted(((int(r[i] or 0), h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h), reverse=True)[:3]
        out.append((s, ex, fname, r[0], r[1].strip()[:95], reasons))
    tot = sum(o[0] for o in out)
    print("total samples", tot, " total instr", sum(o[1] for o in out))
    for s, ex, f, ln, src, rs in sorted(out, reverse=True)[:top]:
        print("%5d %5.1f%% ex %9d %-15s:%-4s %s   %s" % (s, 100.0 * s / max(tot, 1), ex, f, ln, src, [x for x in rs if x[0] > 0]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
