"""Turn what a gpurun ncu call brought back (gpurun_out/) into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_c2.csv            > profiles/rNN_launches_c2.txt
    python profiles/summarize.py metrics  gpurun_out/prof_c2.ncu-rep            > profiles/rNN_metrics_c2.txt
    python profiles/summarize.py opcodes  gpurun_out/prof_c2.ncu-rep <kernel regex> [skip] > profiles/rNN_opcodes_<k>.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
           "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us (ncu: cold cache, serialised -- compare shares)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1]:11.1f} us  x{a[0]:<4d} {100 * a[1] / tot:5.1f}%  {k}")


def _raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def metrics(rep):
    rows = _raw(rep)
    hdr, units = rows[0], rows[1]
    idx = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[ki][:100]}")
        for m, i in idx:
            print(f"   {m:62s} {r[i]:>18s} {units[i]}")


def opcodes(rep, kernel, skip="0"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}",
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    si, ii, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ops, stalls, tot = collections.Counter(), collections.Counter(), 0
    for r in data:
        try:
            n, s = int(r[ii]), int(r[st])
        except (ValueError, IndexError):
            continue
        toks = r[si].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += n
        stalls[op] += s
        tot += n
    print(f"# {rows[0][1][:120]}\n# warp instructions executed: {tot}")
    for op, n in ops.most_common(30):
        print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%   stall samples {stalls[op]}")


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics, "opcodes": opcodes}[sys.argv[1]](*sys.argv[2:])
