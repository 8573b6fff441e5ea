#!/usr/bin/env python
"""Summarise an Nsight Compute report (ncu --set full) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]

Prints per captured launch: duration, DRAM bytes (read+write), DRAM/SM throughput %, occupancy,
registers, the warp-stall mix, and the 25 hottest SASS lines by stall samples with their source.
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        # memory pipeline of the SM: is the kernel bound by L1TEX wavefronts (uncoalesced gathers / scatters)?
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__lsuin_requests.sum",
        "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__warps_issue_stalled_lg_throttle.avg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {rep}: {len(data)} captured launches", file=out)
    for r in data:
        print(f"\n## {r[name_i][:100]}", file=out)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {r[i]:>16s} {units[i]}", file=out)
        stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        if not stalls:
            stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith(".pct")]
        for v, h in sorted(stalls, reverse=True)[:12]:
            print(f"  stall {h:70s} {v:10.3f}", file=out)
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    heads = [i for i, r in enumerate(src) if r and r[0] in ("Address", "#")]
    if heads:
        h = src[heads[0]]
        end = heads[1] - 2 if len(heads) > 1 else len(src)
        body = [r for r in src[heads[0] + 1:end] if len(r) == len(h)]
        si = h.index("# Samples") if "# Samples" in h else None
        ci = h.index("Instructions Executed")
        srci = h.index("Source")
        if si is not None:
            tot = sum(int(r[si] or 0) for r in body)
            print(f"\n## hottest SASS of the first captured launch (of {tot} stall samples, "
                  f"{sum(int(r[ci] or 0) for r in body)} warp instructions)", file=out)
            stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            for r in sorted(body, key=lambda r: -int(r[si] or 0))[:25]:
                st = sorted([(int(r[i] or 0), h[i]) for i in stall_cols], reverse=True)[:2]
                print(f"{int(r[si] or 0):6d} {r[ci]:>9s}  {r[srci][:70]:70s} {st}", file=out)


if __name__ == "__main__":
    main()
