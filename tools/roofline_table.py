"""Per-stage roofline table (markdown) from a bench JSON line: time, algorithmic TFLOP/s and GB/s of every
kernel group against the measured peaks.

    python tools/roofline_table.py profiles/r1_bench_tf32.json
"""
import json, os, sys

d = json.load(open(sys.argv[1]))
r = d["roofline"]
peak_t = r["peak"]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    hbm = float(pk.get("hbm_gbs_sustained") or pk.get("hbm_gbs") or 6548.2)
except Exception:
    hbm = 6548.2
groups = {}
order = []
for p in r["per_stage"]:
    g = p["kernel"].split(".")[0]
    if g not in groups:
        groups[g] = [0.0, 0.0, 0.0, 0]
        order.append(g)
    ms = p["ms"]
    groups[g][0] += ms
    groups[g][1] += (p["tflops"] or 0.0) * ms
    groups[g][2] += (p["gbs"] or 0.0) * ms
    groups[g][3] += p["launches"]
tot = sum(v[0] for v in groups.values())
print(f"mode {d['dtype']}: tensor peak used {peak_t:.0f} TFLOP/s, HBM peak used {hbm:.0f} GB/s, serial sum {tot:.3f} ms\n")
print("| kernel group | launches | ms | share | TFLOP/s | % tensor peak | GB/s (algorithmic) | % HBM peak | bound |")
print("|---|---|---|---|---|---|---|---|---|")
for g in order:
    ms, tf, gb, n = groups[g]
    tf, gb = tf / ms, gb / ms
    ft, fh = tf / peak_t, gb / hbm
    bound = "tensor" if ft >= fh else "hbm"
    print(f"| {g} | {n} | {ms:.3f} | {ms / tot:.3f} | {tf:.0f} | {100 * ft:.0f} % | {gb:.0f} | {100 * fh:.0f} % | {bound} |")
