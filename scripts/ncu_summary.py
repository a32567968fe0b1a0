#!/usr/bin/env python
"""Turn the ncu captures under gpurun_out/ into the small, tracked summaries under profiles/.

    python scripts/ncu_summary.py r01            # writes profiles/r01_*.{md,csv,json}

Inputs (written by scripts/gpu_check.sh on the GPU box):
    gpurun_out/launches.csv              ncu --metrics gpu__time_duration.sum launch list of `python bench.py`
    gpurun_out/prof_<kernel>.ncu-rep     one `ncu --set full --import-source on` capture per hot kernel
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_metrics(rep):
    rows = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    if len(rows) < 3:
        return None
    hdr, units, row = rows[0], rows[1], rows[2]
    d = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
    return d


def source_mix(rep):
    rows = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]).splitlines()))
    if len(rows) < 3:
        return None, None
    hdr = rows[1]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    mix, hot = collections.Counter(), []
    for idx, r in enumerate(rows[2:]):
        if len(r) <= i_ex or not r[i_ex].isdigit():
            continue
        op = r[i_src].strip().split()
        if not op:
            continue
        name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        mix[name.split(".")[0].rstrip(";")] += int(r[i_ex])
        hot.append((int(r[i_s]) if r[i_s].isdigit() else 0, idx, r[i_src].strip()))
    return mix, sorted(hot, reverse=True)[:15]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    summary = {}
    for kern in ("k_step", "k_raster_follow", "k_forces", "k_rasterize"):
        rep = os.path.join(OUT, "prof_%s.ncu-rep" % kern)
        if not os.path.isfile(rep):
            continue
        d = raw_metrics(rep)
        if d is None:
            continue
        mix, hot = source_mix(rep)
        lines = ["# ncu --set full: %s  (%s)" % (kern, d.get("Kernel Name", ("?", ""))[0]), "",
                 "command: `ncu --set full --clock-control none --import-source on -k regex:%s --launch-skip 4 -c 1 "
                 "python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --no-paac` (C4: 4096 envs x 256 locusts; ncu "
                 "serialises kernels, so k_step and k_raster_follow -- concurrent in production -- are each seen alone)" % kern, "",
                 "| metric | value | unit |", "|---|---|---|"]
        vals = {}
        for k in KEYS:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
                vals[k] = d[k][0]
        if mix:
            tot = sum(mix.values())
            lines += ["", "## executed warp-instruction mix (SASS opcode, share of %d)" % tot, "",
                      "| opcode | warp instr | share |", "|---|---|---|"]
            for op, n in mix.most_common(18):
                lines.append("| %s | %d | %.1f %% |" % (op, n, 100.0 * n / tot))
            vals["mix"] = dict(mix.most_common(18))
        if hot:
            lines += ["", "## top stall-sample instructions", "", "| samples | sass line | instruction |", "|---|---|---|"]
            for sm, idx, src in hot:
                lines.append("| %d | %d | `%s` |" % (sm, idx, src))
        with open(os.path.join(PROF, "%s_ncu_%s.md" % (tag, kern)), "w") as f:
            f.write("\n".join(lines) + "\n")
        summary[kern] = vals
    # launch list
    ll = os.path.join(OUT, "launches.csv")
    if os.path.isfile(ll):
        rows = [r for r in csv.reader(open(ll)) if len(r) > 14 and r[0].isdigit()]
        per = collections.OrderedDict()
        for r in rows:
            name = r[4].split("(")[0].replace("void ", "").strip()[:90]
            t = float(r[-1])
            a = per.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += t
        tot = sum(v[1] for v in per.values())
        with open(os.path.join(PROF, "%s_launches.md" % tag), "w") as f:
            f.write("# ncu launch list of `python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph` (C4)\n\n"
                    "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` -- %d launches, %.1f us in total "
                    "(cold-cache, serialised: shares matter, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n"
                    % (len(rows), tot / 1e3))
            for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
                f.write("| %s | %d | %.1f | %.1f %% |\n" % (name, n, t / 1e3, 100.0 * t / tot))
        summary["launches"] = {k: {"n": v[0], "us": v[1] / 1e3} for k, v in per.items()}
        import shutil
        shutil.copy(ll, os.path.join(PROF, "%s_launches.csv" % tag))
    with open(os.path.join(PROF, "%s_summary.json" % tag), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
