"""Per-kernel-family summary of ONE captured step from an ncu launch list (csv with gpu__time_duration.sum and, if
present, dram__bytes_read.sum / dram__bytes_write.sum).  The step is the last run of es:: kernels that ends at the
first cfg_ddim launch (the first CUDA-graph replay in bench.py).
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
by_id = collections.OrderedDict()
for r in rows:
    k = r["ID"]
    d = by_id.setdefault(k, {"name": r["Kernel Name"], "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m.startswith("gpu__time_duration"):
        d["us"] = v / 1e3 if u.startswith("ns") else (v * 1e3 if u.startswith("ms") else v)
    else:
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["rd" if "read" in m else "wr"] = v * mult
launches = list(by_id.values())
end = next((i for i, l in enumerate(launches) if "cfg_ddim" in l["name"]), len(launches))
start = max(i for i, l in enumerate(launches[:end]) if "timestep_embedding" in l["name"])
while start > 0 and ("nchw_to_nhwc" in launches[start - 1]["name"] or "im2col" in launches[start - 1]["name"]):
    start -= 1
step = launches[start:end + 1]
fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
for l in step:
    n = re.sub(r"<.*", "", l["name"]).replace("void ", "")
    n = re.sub(r"\(.*", "", n)
    f = fam[n]
    f[0] += 1
    f[1] += l["us"]
    f[2] += l["rd"] + l["wr"]
tot = sum(f[1] for f in fam.values())
print("| kernel | launches/step | sum of durations (us) | share | DRAM bytes (MB) |")
print("|---|---|---|---|---|")
for n, (c, us, by) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {us:.0f} | {100 * us / tot:.1f}% | {by / 1e6:.0f} |")
print()
print(f"launches/step {len(step)}; sum of per-launch durations {tot / 1e3:.2f} ms (cold-cache, serialised by ncu); "
      f"DRAM read {sum(l['rd'] for l in step) / 1e9:.2f} GB + write {sum(l['wr'] for l in step) / 1e9:.2f} GB per step")
# machine-readable copy for bench.py (roofline.traffic is read from this file, never hard-coded there)
if len(sys.argv) > 2:
    import json

    json.dump({"dram_bytes_per_step": sum(l["rd"] + l["wr"] for l in step), "launches_per_step": len(step),
               "sum_of_durations_us": tot, "source": f"{sys.argv[1]} via tools/launch_summary.py (ncu, caches flushed per launch)",
               "families": {n: {"launches": c, "us": round(us, 1), "dram_mb": round(by / 1e6, 1)} for n, (c, us, by) in fam.items()}},
              open(sys.argv[2], "w"), indent=1)
