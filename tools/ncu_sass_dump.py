#!/usr/bin/env python
"""Per-SASS-instruction table of one kernel in an ncu report, as a small csv (address, executed warp-instructions,
thread-instructions, stall samples, SASS text) -- to be read next to `nvdisasm -g` of the same cubin.

    python tools/ncu_sass_dump.py <report.ncu-rep> <out.csv.gz>
"""
import csv
import gzip
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1:3]
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"],
                                                      capture_output=True, text=True).stdout)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[heads[0]]
    end = heads[1] - 1 if len(heads) > 1 else len(rows)
    body = [r for r in rows[heads[0] + 1:end] if len(r) == len(h)]
    if "--all" in sys.argv:                       # every column of the page (memory transactions per instruction, stall reasons)
        with gzip.open(out, "wt", newline="") as f:
            w = csv.writer(f)
            w.writerow(h)
            w.writerows(body)
        print(len(body), "rows,", len(h), "columns")
        return
    cols = [h.index(c) for c in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples", "Source")]
    with gzip.open(out, "wt", newline="") as f:
        w = csv.writer(f)
        w.writerow(["address", "inst", "thread_inst", "samples", "sass"])
        for r in body:
            w.writerow([r[c] for c in cols])
    print(len(body), "rows")


if __name__ == "__main__":
    main()
