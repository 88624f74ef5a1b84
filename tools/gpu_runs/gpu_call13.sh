#!/bin/bash
out=gpurun_out/c13
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -s -k "other_geometries or rejects_bad or other_configurations or odd_upsample_geometry" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "passed|failed|max-abs|Error|assert" $out/pytest.log | head -30
