#!/usr/bin/env python
"""One ncu --set full capture -> a small tracked summary:  python scripts/ncu_report.py REP OUT.md "what was run"
Reuses the metric list / source-page reader of scripts/ncu_summary.py; also writes OUT.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_summary as ns


def main():
    rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    d = ns.raw_metrics(rep)
    mix, hot = ns.source_mix(rep)
    lines = ["# ncu --set full: %s" % d.get("Kernel Name", ("?", ""))[0], "", what, "", "| metric | value | unit |", "|---|---|---|"]
    vals = {}
    for k in ns.KEYS:
        if k in d:
            lines.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
            vals[k] = d[k][0]
            vals[k + "__unit"] = d[k][1]
    if mix:
        tot = sum(mix.values())
        lines += ["", "## executed warp-instruction mix (SASS opcode, share of %d)" % tot, "", "| opcode | warp instr | share |",
                  "|---|---|---|"]
        for op, n in mix.most_common(18):
            lines.append("| %s | %d | %.1f %% |" % (op, n, 100.0 * n / tot))
        vals["mix"] = dict(mix.most_common(18))
    if hot:
        lines += ["", "## top stall-sample instructions", "", "| samples | sass line | instruction |", "|---|---|---|"]
        for sm, idx, src in hot:
            lines.append("| %d | %d | `%s` |" % (sm, idx, src))
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.splitext(out)[0] + ".json", "w") as f:
        json.dump(vals, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
