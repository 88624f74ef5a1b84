#!/bin/bash
out=gpurun_out/c21
mkdir -p $out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "unusual_resblock" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt
grep -E "structure|passed|failed|Error|assert" $out/pytest.log | tail -20
