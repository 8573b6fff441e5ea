#!/usr/bin/env python
"""Attribute executed warp-instructions / stall samples of an ncu report to SOURCE LINES.

    python tools/ncu_lines.py <report.ncu-rep> <libev2b.so> <mangled-kernel-substring> [top]

ncu's csv SASS page has no line column, so the line table comes from `nvdisasm -g` of the same
cubin; the two listings are aligned by instruction order within the kernel.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines, cur, on = [], None, False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            lines.append(cur)
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"],
                                                      capture_output=True, text=True).stdout)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[heads[0]]
    end = heads[1] - 1 if len(heads) > 1 else len(rows)
    body = [r for r in rows[heads[0] + 1:end] if len(r) == len(h)]
    ci, si, ti = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    if len(body) != len(lines):
        print(f"warning: {len(body)} SASS rows in the report vs {len(lines)} in nvdisasm; aligning the prefix")
    inst, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r, l in zip(body, lines):
        inst[l] += int(r[ci] or 0); smp[l] += int(r[si] or 0); thr[l] += int(r[ti] or 0)
    srcdir = os.path.dirname(os.path.abspath(so))
    srcs = {f: open(os.path.join(srcdir, f)).read().splitlines() for f in ("ev2b_device.cuh", "ev2b_evlist.cuh", "ev2b_math.h", "ev2b.cu")
            if os.path.exists(os.path.join(srcdir, f))}
    ti_, ts_ = sum(inst.values()), sum(smp.values())
    print(f"total warp-instructions {ti_}, stall samples {ts_}")
    print("  file:line               inst%  smp%  thr/inst  source")
    for l, n in inst.most_common(top):
        f, no = l if l else ("?", 0)
        src = srcs.get(f)
        text = src[no - 1].strip()[:90] if src and 0 < no <= len(src) else "(CUDA header)"
        print(f"{f[:16] + ':' + str(no):>22} {100 * n / ti_:6.2f} {100 * smp[l] / max(ts_, 1):5.1f} {thr[l] / max(n, 1):8.1f}  {text}")


if __name__ == "__main__":
    main()
