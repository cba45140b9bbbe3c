#!/usr/bin/env python3
"""Turn an ncu report (.ncu-rep) into the small text summaries kept under profiles/.
Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/<name>   (run here, no GPU needed)"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct",
        "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "sm__inst_executed_pipe_lsu",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "smsp__thread_inst_executed_per_inst_executed",
        "sm__throughput.avg.pct", "smsp__warp_issue_stalled", "smsp__average_warps_issue_stalled",
        "sm__pipe_tensor_cycles_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out + "_raw.txt", "w") as fh:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            fh.write("== kernel: %s\n" % name[:160])
            for i, h in enumerate(hdr):
                if any(h.startswith(k) for k in KEYS):
                    fh.write("%-90s %-14s %s\n" % (h, units[i], r[i]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data = rows[2:]
    tot = sum(int(r[iE]) for r in data)
    ops = {}
    for r in data:
        op = r[iS].split()[0].split(".")[0] if r[iS].split() else "?"
        if op.startswith("@"):
            op = r[iS].split()[1].split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[iE])
    with open(out + "_source.txt", "w") as fh:
        fh.write("total warp instructions executed: %d\n\ninstruction mix (executed):\n" % tot)
        for op, n in sorted(ops.items(), key=lambda t: -t[1])[:25]:
            fh.write("  %-12s %12d  %5.1f%%\n" % (op, n, 100.0 * n / tot))
        fh.write("\nhottest SASS lines (by stall samples):\n")
        for r in sorted(data, key=lambda r: -int(r[iN]))[:40]:
            fh.write("  %6s samples  %10s exec  %s\n" % (r[iN], r[iE], r[iS].strip()[:90]))
    print("wrote", out + "_raw.txt", out + "_source.txt")


if __name__ == "__main__":
    main()
