"""Phase timeline of one tc_conv_kernel launch (tuning only).
    HFG_TC_CONV_TIMELINE=ups3:gpurun_out/ctl.txt python tools/profile_step.py --mode tf32 --forwards 1
    python tools/conv_timeline.py gpurun_out/ctl.txt
Events per CTA: 0 entry, 1 setup done, 2 first activations landed (MMA warp), 3 all MMAs issued,
4 accumulator complete (epilogue warp), 5 epilogue done, 6 exit."""
import sys
for block in open(sys.argv[1]).read().split("#")[1:]:
    lines = block.strip().split("\n")
    rows = [[int(x) for x in l.split()] for l in lines[1:] if l.strip()]
    rows = [r for r in rows if r[6]]
    print(lines[0])
    n = len(rows)
    avg = lambda f: sum(f(r) for r in rows) / n
    print(f"  CTAs stamped {n}: setup {avg(lambda r: r[1]-r[0]):.0f} | wait first A {avg(lambda r: r[2]-r[1]):.0f} | "
          f"MMA issue {avg(lambda r: r[3]-r[2]):.0f} | issue end -> acc complete {avg(lambda r: r[4]-r[3]):.0f} | "
          f"epilogue {avg(lambda r: r[5]-r[4]):.0f} | exit {avg(lambda r: r[6]-r[5]):.0f} | total {avg(lambda r: r[6]-r[0]):.0f} cycles")
