#!/bin/bash
# gpurun_out/r2 ncu captures -> tracked summaries under profiles/ (run here, after scripts/gpu_r2_profile.sh)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2
what() { echo "\`ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip 6 -c 1 python scripts/run_steps.py $3 $4\` ($3 envs x $4 locusts; ncu serialises kernels: a kernel that is concurrent with another in production is seen alone)"; }
for spec in "k_step_c4 k_step 4096 256" "k_raster_follow_c4 k_raster_follow 4096 256" "k_step_512x256 k_step 512 256" "k_step_n80 k_step 1024 80" "k_step_c2 k_step 1024 64" "k_forces_c4 k_forces 4096 256"; do
  set -- $spec
  [ -f $O/prof_r02_$1.ncu-rep ] && python scripts/ncu_report.py $O/prof_r02_$1.ncu-rep profiles/r02_ncu_$1.md "$(what $@)" && echo "profiles/r02_ncu_$1.md"
done
python - <<'PY'
import csv, collections, json, os
O = "gpurun_out/r2"
summ = {}
for f in sorted(os.listdir("profiles")):
    if f.startswith("r02_ncu_") and f.endswith(".json"):
        summ[f[len("r02_ncu_"):-5]] = json.load(open(os.path.join("profiles", f)))
def traffic(*names):
    t = 0.0
    for n in names:
        d = summ.get(n, {})
        if "dram__bytes_read.sum" in d:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                t += float(d[m]) * mult.get(d.get(m + "__unit", "Mbyte"), 1e6)
    return t or None
summ["traffic_bytes_per_step"] = {"c4": traffic("k_step_c4", "k_raster_follow_c4"), "c2": traffic("k_step_c2"), "n80": traffic("k_step_n80")}
ll = os.path.join(O, "launches_r02.csv")
if os.path.isfile(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 14 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").strip()[:100]
        a = per.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
    tot = sum(v[1] for v in per.values())
    with open("profiles/r02_launches.md", "w") as f:
        f.write("# ncu launch list of `python bench.py --steps 8 --warmup 3 --no-graph --no-cpu-baseline --no-paac --no-secondary` (C4)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` -- first %d launches, %.1f us in total "
                "(cold-cache, serialised: shares matter, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % (len(rows), tot / 1e3))
        for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (name, n, t / 1e3, 100.0 * t / tot))
    import shutil; shutil.copy(ll, "profiles/r02_launches.csv")
    summ["launches"] = {k: {"n": v[0], "us": v[1] / 1e3} for k, v in per.items()}
json.dump(summ, open("profiles/r02_summary.json", "w"), indent=1, sort_keys=True)
print("traffic", summ["traffic_bytes_per_step"])
PY
