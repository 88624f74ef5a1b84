#!/bin/bash
out=gpurun_out/c17
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt
tail -5 $out/pytest_gpu.log
