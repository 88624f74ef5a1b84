#!/bin/bash
# final verification of the shipped code: smoke(), the parity subset that touches every kernel variant, the bit-identity groups
out=gpurun_out/f3
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee $out/rc.txt
grep smoke $out/smoke.log
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or config2 or variants or unusual or fp32_plane or stage_boundaries or saturated" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -3 $out/pytest.log
