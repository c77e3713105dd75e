"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches <out.md> <launches.csv> "<command that was profiled>"
  python tools/summarize_ncu.py kernel   <out.md> <prof.ncu-rep> "<title>" "<workload>" [<bench.json> [<reference.json>]]
"""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

mode, out = sys.argv[1], Path(sys.argv[2])
out.parent.mkdir(exist_ok=True)

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]

if mode == "launches":
    launches, cmd = sys.argv[3], sys.argv[4]
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per, lines = collections.OrderedDict(), []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])[:90]
        t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        per.setdefault(name, []).append(t)
        lines.append((name, t))
    total = sum(t for _, t in lines)
    with open(out, "w") as f:
        f.write(f"# launch list of `{cmd}` under `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n"
                "Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes.\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, ts in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{name}` | {len(ts)} | {sum(ts):.3f} | {100 * sum(ts) / total:.2f}% |\n")
        solve = [t for n, t in lines if "solve_" in n]
        f.write(f"\nsolver launches (ms): {', '.join(f'{t:.1f}' for t in solve)}\n")
        prep = sum(sum(ts) for n, ts in per.items() if "cub" in n or "work_keys" in n) / max(len(solve), 1)
        f.write("\nThe DFMA peak probe (`dfma_peak_kernel`) runs before the timed region; within one solver pass "
                "(work_keys + the CUB radix-sort kernels + the solver kernel) the solver kernel is "
                f"{100 * max(solve) / (max(solve) + prep):.3f}% of the device time.  Since the latency lane became conditional, each "
                "pass enqueues BOTH `duo_solve_kernel` (latency lane in front of the one-set-per-warp queue) and the plain "
                "`solve_kernel`; a guard word written by the plan kernel lets exactly one of them run (the 0.0 ms entries are the "
                "other one returning at once).\n")
    print("wrote", out)
else:
    rep, title, workload = sys.argv[3], sys.argv[4], sys.argv[5]
    bench = sys.argv[6] if len(sys.argv) > 6 else None
    ref = sys.argv[7] if len(sys.argv) > 7 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    m = dict(zip(rr[0], zip(rr[1], rr[2])))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sr = list(csv.reader(src.splitlines()))
    sh = sr[1]
    isrc, ie = sh.index("Source"), sh.index("Instructions Executed")
    mix = collections.Counter()
    for r in sr[2:]:
        try:
            ex = int(r[ie])
        except (ValueError, IndexError):
            continue
        mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        op = mm.group(2) if mm else "?"
        mix[op if "MOV" in op else op.split(".")[0]] += ex
    with open(out, "w") as f:
        f.write(f"# `ncu --set full --clock-control none` of {title}\n\nKernel: `{rr[2][rr[0].index('Kernel Name')] if 'Kernel Name' in rr[0] else ''}`\n\n"
                f"Workload: {workload}\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in KEYS:
            if k in m:
                f.write(f"| {k} | {m[k][1]} | {m[k][0]} |\n")
        tot = sum(mix.values())
        f.write(f"\n## Executed warp-instructions by opcode (whole launch, {tot:.4g} total)\n\n| opcode | count | share |\n|---|---:|---:|\n")
        for op, c in mix.most_common(20):
            f.write(f"| {op} | {c:.4g} | {100 * c / tot:.1f}% |\n")
        dp = sum(c for op, c in mix.items() if op in ("DFMA", "DMUL", "DADD", "DSETP"))
        f.write(f"\nFP64-pipe instructions (DFMA+DMUL+DADD+DSETP): {dp:.4g} = {100 * dp / tot:.1f}% of all issued warp-instructions.\n")
        if bench:
            b = json.loads([l for l in open(bench) if l.startswith("{")][-1])
            f.write("\n## Bench line this capture belongs to (taken WITHOUT ncu in the same gpurun call)\n\n```json\n" + json.dumps(b, indent=1) + "\n```\n")
        if ref:
            f.write("\n## Reference arm (`bench.py --impl reference`, same box)\n\n```json\n" +
                    json.dumps(json.loads([l for l in open(ref) if l.startswith("{")][-1]), indent=1) + "\n```\n")
    print("wrote", out)
