"""Hot source lines of an ncu capture taken with --import-source on (warp-stall samples per CUDA-C line).

    python tools/ncu_source_hot.py gpurun_out/c16/pair_mrf1k3_tf32lo.ncu-rep [N lines]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]


def toint(v):
    try:
        return int(v)
    except ValueError:
        return 0


agg, inst, src = collections.Counter(), collections.Counter(), {}
stall = collections.defaultdict(collections.Counter)
for hi, i in enumerate(hdr_idx):
    hdr, fpath = rows[i], rows[i - 2][1]
    end = hdr_idx[hi + 1] - 2 if hi + 1 < len(hdr_idx) else len(rows)
    cols = {}
    for j, n in enumerate(hdr):
        cols.setdefault(n, j)
    for r in rows[i + 1:end]:
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        key = (fpath.split("/")[-1], int(r[0]))
        agg[key] += toint(r[cols["# Samples"]])
        inst[key] += toint(r[cols["Instructions Executed"]])
        src[key] = r[1]
        for n in hdr:
            if n.startswith("stall_") and "Not Issued" not in n:
                v = toint(r[cols[n]])
                if v:
                    stall[key][n] += v
tot = sum(agg.values())
print(f"{rep}: {tot} samples, {sum(inst.values())} warp instructions")
allst = collections.Counter()
for k in stall:
    allst.update(stall[k])
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in allst.most_common(8)))
for key, s in agg.most_common(top):
    t3 = ", ".join(f"{k[6:]}:{v}" for k, v in stall[key].most_common(3))
    print(f"{100 * s / tot:5.1f}% inst {inst[key]:8d} {key[0][:20]}:{key[1]:4d} {src[key].strip()[:86]}  [{t3}]")
