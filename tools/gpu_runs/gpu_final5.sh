#!/bin/bash
out=gpurun_out/f5
mkdir -p $out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "arithmetic_model_predicts" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt
grep -E "max-abs vs reference|passed|failed|assert" $out/pytest.log | tail -12
