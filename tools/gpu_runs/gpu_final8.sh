#!/bin/bash
# the plain-C host example against the Python mirror and the goldens
out=gpurun_out/f8
mkdir -p $out
timeout 55 python -m pytest tests/test_zz_c_host.py -m gpu -q -s -p no:cacheprovider > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt
grep -E "c_host|passed|failed|Error|assert" $out/pytest.log | tail -12
