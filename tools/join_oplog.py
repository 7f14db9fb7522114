"""Join the op log of tools/step_profile.py with an ncu launch list of the same run (es:: kernels in order)."""
import collections, csv, json, re, sys
oplog = json.load(open(sys.argv[1]))
with open(sys.argv[2]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if "es::" in r["Kernel Name"]]
def us(row):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    return v / 1e3 if u.startswith("ns") else (v * 1e3 if u.startswith("ms") else v)
nk = {"gemm": 1, "attention": 1, "groupnorm": 2, "layernorm": 1, "merge": 3, "small_linear": 1, "im2col3x3": 1,
      "upsample2x": 1, "nchw_to_nhwc": 1, "timestep_embedding": 1}
need = sum(o[4] for o in oplog)
rows = rows[-need:]
i = 0
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
fam = collections.defaultdict(float)
for name, desc, fl, nb, nl in oplog:
    t = sum(us(r) for r in rows[i:i + nl]); i += nl
    k = (name, desc); by[k][0] += 1; by[k][1] += t; by[k][2] += fl; by[k][3] += nb; fam[name] += t
print("total us", sum(fam.values()), {k: round(v) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})
print(f"{'us total':>9s} {'n':>3s} {'us avg':>8s} {'TFLOP/s':>8s} {'GB/s':>7s}  op")
for (name, desc), (n, t, fl, nb) in sorted(by.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 60]:
    print(f"{t:9.1f} {n:3d} {t / n:8.1f} {fl / t / 1e6:8.1f} {nb / t / 1e3:7.0f}  {name} {desc}")
