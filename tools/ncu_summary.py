#!/usr/bin/env python
"""Text summary of an .ncu-rep (read here, without a GPU): per launch the headline counters, then per kernel
the SASS regions by executed-instruction share (needs -lineinfo / --import-source on).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [units_per_launch] > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return [r for r in csv.reader(io.StringIO(out))]


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = page(rep, "raw")
    hdr, unit_row = raw[0], raw[1]
    print("# %s" % rep)
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, unit_row))
        print("\n## launch %s  %s" % (d.get("ID"), d.get("Kernel Name")))
        for k in KEYS:
            if k in d and d[k] != "":
                print("  %-72s %s %s" % (k, d[k], u.get(k, "")))
        stalls = sorted(((float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")),
                        reverse=True)
        print("  stall reasons (warps per issue-active cycle): " + ", ".join(
            "%s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], v) for v, k in stalls[:7]))
        if units:
            print("  warp-instructions per unit (%g units/launch): %.1f" % (units, float(d["smsp__inst_executed.sum"]) / units))
    src = page(rep, "source")
    # the source page concatenates kernels: header rows start with "Kernel Name"
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        rows = b["rows"]
        if len(rows) < 2:
            continue
        h = rows[0]
        ix = {n: i for i, n in enumerate(h)}
        data = rows[1:]
        if "Instructions Executed" not in ix:
            continue
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
        samp = sum(int(r[ix["# Samples"]]) for r in data) or 1
        print("\n## SASS regions of %s (%d instructions executed, %d samples)" % (b["name"], tot, samp))
        print("  %-12s %5s %14s %8s %8s  first instruction" % ("sass idx", "len", "exec/instr", "inst %", "smpl %"))
        prev, start, acc, cnt, sacc = None, 0, 0, 0, 0

        def flush(k):
            if prev is not None and (acc > 0.01 * tot or sacc > 0.01 * samp):
                print("  %5d-%-6d %5d %14.0f %7.1f%% %7.1f%%  %s" % (start, k - 1, cnt, prev, 100.0 * acc / tot,
                                                                 100.0 * sacc / samp, data[start][ix["Source"]].strip()[:60]))
        for k, r in enumerate(data):
            e = int(r[ix["Instructions Executed"]])
            if prev is None or abs(e - prev) > max(0.0004 * tot / 10, 0.15 * prev):
                flush(k)
                start, acc, cnt, sacc, prev = k, 0, 0, 0, e
            acc += e
            cnt += 1
            sacc += int(r[ix["# Samples"]])
        flush(len(data))


if __name__ == "__main__":
    main()
