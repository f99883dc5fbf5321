#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 > gpurun_out/tests_f_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_f_conv.log
timeout 600 python bench.py --steps 3 --warmup 2 --engine native --no-cpu-baseline > gpurun_out/bench_f_native.json 2> gpurun_out/bench_f_native.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_f_native.csv python bench.py --steps 1 --warmup 1 --clips 2 --engine native --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1
grep -E "passed|failed" gpurun_out/tests_f_conv.log | tail -2; grep -E "^(FAILED|E  )" gpurun_out/tests_f_conv.log | head -30 | cut -c1-220; head -c 400 gpurun_out/bench_f_native.json; tail -3 gpurun_out/bench_f_native.err
