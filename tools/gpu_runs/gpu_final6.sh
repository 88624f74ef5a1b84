#!/bin/bash
out=gpurun_out/f6
mkdir -p $out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "variants and tf32" --durations=3 > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt
tail -8 $out/pytest.log
