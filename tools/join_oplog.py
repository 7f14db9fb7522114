"""Joins an ncu launch list of an EAGER `tools/step_profile.py` run (launches appear in host issue order) with the op log that
run wrote (`gpurun_out/oplog.json`: name, shape, flops, bytes, launches per op) -> per-shape kernel time without the host
launch latency that CUDA events around a single eager call include.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eager.csv python tools/step_profile.py 1
    python tools/join_oplog.py gpurun_out/eager.csv gpurun_out/oplog.json > profiles/<name>.txt"""
import collections
import csv
import json
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"].startswith("gpu__time_duration")]
es = []
for r in rows:
    if "es::" not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    es.append((r["Kernel Name"], v / 1e3 if u.startswith("ns") else (v * 1e3 if u.startswith("ms") else v), r["Grid Size"]))
ops = json.load(open(sys.argv[2]))
need = sum(o[4] for o in ops)
assert len(es) >= need, (len(es), need)
es = es[-need:]
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, set()])
i = 0
for name, desc, fl, nb, nl in ops:
    us = sum(e[1] for e in es[i:i + nl])
    kern = {e[0].split("(")[0].replace("void es::", "").split("<")[0] + " " + e[2] for e in es[i:i + nl]}
    i += nl
    b = by[(name, desc)]
    b[0] += 1
    b[1] += us
    b[2] += fl
    b[3] += nb
    b[4] |= kern
tot = sum(b[1] for b in by.values())
fam = collections.defaultdict(float)
for (name, _), b in by.items():
    fam[name] += b[1]
print(f"ops {len(ops)}  launches {need}  sum of kernel durations {tot / 1e3:.2f} ms (ncu, serialised, cold-ish)")
print({k: round(v / 1e3, 2) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})
print(f"{'us total':>9s} {'n':>3s} {'us avg':>8s} {'TFLOP/s':>8s} {'GB/s':>7s}  op  [kernels grid]")
for (name, desc), (n, us, fl, nb, kern) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.1f} {n:3d} {us / n:8.1f} {fl / us / 1e6:8.1f} {nb / us / 1e3:7.0f}  {name} {desc}  {sorted(kern)[:3]}")
