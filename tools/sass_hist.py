#!/usr/bin/env python
"""Instruction-mix / stall histogram from `ncu -i rep --page source --csv --kernel-name regex:X > file.csv`."""
import collections
import csv
import sys


def main(path, top=22):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(hdr)}
    ops, samples, execd = collections.Counter(), collections.Counter(), collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in rows:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        toks = r[col["Source"]].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        ops[op] += 1
        samples[op] += int(r[col["# Samples"]] or 0)
        execd[op] += int(r[col["Instructions Executed"]] or 0)
        for s in stall_cols:
            tot[s] += int(r[col[s]] or 0)
    print("static instructions: %d (%.0f KB)   executed: %d   samples: %d" % (sum(ops.values()), sum(ops.values()) * 16 / 1024, sum(execd.values()), sum(samples.values())))
    for op, c in ops.most_common(top):
        print("  %-10s static %6d  executed %10d  samples %7d" % (op, c, execd[op], samples[op]))
    print("stall samples:", ", ".join("%s=%d" % kv for kv in tot.most_common(12)))


if __name__ == "__main__":
    main(sys.argv[1])
