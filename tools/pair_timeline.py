"""Phase timeline of the fused ResBlock-pair kernel (tuning only).

    HFG_TC_TIMELINE=gpurun_out/tl.txt python tools/tune_layers.py --which 2 --stages 3 --mode bf16
    python tools/pair_timeline.py gpurun_out/tl.txt

The kernel stamps clock64 at 12 events per tile for CTAs 0..3 (tc_pair_kernel.cuh, HFG_TL).  This prints,
per layer, the cycles between consecutive events averaged over steady-state tiles (2..) of CTA 0."""
import sys

EV = ["prodA", "mmaA", "c1done", "mmaH", "c2done", "epiA", "pre2", "acc1", "epi1", "acc2", "epi2", "prodW"]
blocks, cur = [], None
for line in open(sys.argv[1]):
    if line.startswith("#"):
        cur = {"hdr": line[1:].strip(), "rows": []}
        blocks.append(cur)
    elif line.strip():
        cur["rows"].append([int(x) for x in line.split()])
for b in blocks:
    print(b["hdr"])
    for cta in range(2):
        tiles = [r for r in b["rows"][cta * 16:(cta + 1) * 16] if r[4] and r[10]]
        if len(tiles) < 4:
            continue
        t0 = tiles[0][1]
        print(f" cta {cta}: tiles stamped {len(tiles)}")
        print("   tile " + " ".join(f"{e:>7s}" for e in EV) + "   (cycles since mmaA of tile 0)")
        for i, r in enumerate(tiles[:8]):
            print(f"   {i:4d} " + " ".join(f"{r[j] - t0:7d}" for j in range(12)))
        st = tiles[2:]
        per = (st[-1][10] - st[0][10]) / max(1, len(st) - 1)
        avg = lambda f: sum(f(r) for r in st) / len(st)
        print(f"   steady state: {per:.0f} cycles/tile | conv1 issue {avg(lambda r: r[2]-r[1]):.0f} | wait H after c1 {avg(lambda r: r[3]-r[2]):.0f} | "
              f"conv2 issue {avg(lambda r: r[4]-r[3]):.0f} | pre2 {avg(lambda r: r[6]-r[5]):.0f} | wait acc1 {avg(lambda r: r[7]-r[6]):.0f} | "
              f"epi1 {avg(lambda r: r[8]-r[7]):.0f} | wait acc2 {avg(lambda r: r[9]-r[8]):.0f} | epi2 {avg(lambda r: r[10]-r[9]):.0f} | "
              f"A issue->mma sees it {avg(lambda r: r[1]-r[0]):.0f}")
