#!/bin/bash
out=gpurun_out/c3
mkdir -p $out
timeout 900 python -m pytest tests/test_ar_decoder.py tests/test_log_mel.py tests/test_length_regulator.py -m gpu -q -s > $out/pytest_new.log 2>&1; echo "pytest_new rc=$?" | tee -a $out/rc.txt
grep -E "passed|failed|error|max-abs|decoder|log-mel|Error" $out/pytest_new.log | head -30
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "golden or saturated or ragged or config2 or chunking" > $out/pytest_par.log 2>&1; echo "pytest_par rc=$?" | tee -a $out/rc.txt
tail -5 $out/pytest_par.log
for m in tf32 bf16; do timeout 200 python tools/stage_times.py $m > $out/stages_$m.txt 2>&1; done
paste $out/stages_tf32.txt $out/stages_bf16.txt
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
tail -3 $out/bench.err
# compute-sanitizer memcheck on a tiny forward of every mode (one tool per call)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py fp32 tf32 fp16 bf16 > $out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $out/rc.txt
tail -8 $out/memcheck.log
