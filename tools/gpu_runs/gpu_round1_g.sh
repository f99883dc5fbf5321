#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_vgg.py vgg 3 32 > gpurun_out/diag_vgg.log 2>&1
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -k "strided or native_engine" > gpurun_out/tests_g_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_g_conv.log
timeout 600 python bench.py --steps 3 --warmup 2 --engine native --no-cpu-baseline > gpurun_out/bench_g_native.json 2> gpurun_out/bench_g_native.err
cat gpurun_out/diag_vgg.log | tail -40; grep -E "passed|failed" gpurun_out/tests_g_conv.log | tail -2; grep -E "^(FAILED|E  )" gpurun_out/tests_g_conv.log | head -20 | cut -c1-220; head -c 300 gpurun_out/bench_g_native.json; tail -3 gpurun_out/bench_g_native.err
