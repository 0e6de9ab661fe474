#!/usr/bin/env python
"""Turns an ncu report (.ncu-rep, captured with --set full --import-source on) or a launch-list csv into the
markdown summaries kept under profiles/.

  python scripts/ncu_summary.py rep  <file.ncu-rep> "<title>" "<command>" > profiles/xxx.md
  python scripts/ncu_summary.py list <launches.csv> "<title>" "<command>" > profiles/yyy.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def rep(path, title, command):
    raw = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
    head, units, vals = raw[0], raw[1], raw[2]
    print(f"# {title}\n\nCommand: `{command}`\n\n| metric | value |\n|---|---|")
    for m in METRICS:
        if m in head:
            i = head.index(m)
            print(f"| {m} | {vals[i]} {units[i]} |")
    rows = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    agg, inst, text = collections.Counter(), collections.Counter(), {}
    fname, hdr = None, None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname, hdr = r[1].split("/")[-1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and fname and r and r[0] not in ("", "Function Name"):
            try:
                ln, s, ie = int(r[0]), int(r[4]), int(r[7])
            except (ValueError, IndexError):
                continue
            k = (fname, ln)
            agg[k] += s
            inst[k] += ie
            text[k] = r[1][:100]
    tot, ti = sum(agg.values()) or 1, sum(inst.values()) or 1
    print(f"\n## warp-stall samples by CUDA source line (top)\n\n```\nsamples {tot} instr {ti}")
    for k, v in agg.most_common(22):
        print(f"{v:7d} {100 * v / tot:5.1f}%  inst {100 * inst[k] / ti:5.1f}%  {k[0]}:{k[1]}  {text[k]}")
    print("```")


def launches(path, title, command):
    tot = collections.defaultdict(lambda: [0, 0.0])
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if not l.startswith("=="))]
    head = rows[0]
    ik, iv = head.index("Kernel Name"), head.index("Metric Value")
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        tot[name][0] += 1
        tot[name][1] += float(r[iv].replace(",", "")) / 1e6
    total = sum(v[1] for v in tot.values())
    print(f"# {title}\n\nCommand: `{command}`\n(per-launch times are cold-cache and serialised: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f}% | {1e3 * ms / n:.1f} |")
    print(f"\nTotal {total:.1f} ms over {sum(v[0] for v in tot.values())} launches.")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](*sys.argv[2:5])
