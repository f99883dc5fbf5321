#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_parity.py resnet cudnn native > gpurun_out/diag_d.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x > gpurun_out/tests_d_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_d_conv.log
timeout 1200 python -m pytest tests/test_gpu_attacks.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/tests_d_rest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_d_rest.log
timeout 600 python bench.py --steps 3 --warmup 2 --engine native --no-cpu-baseline > gpurun_out/bench_native_tc.json 2> gpurun_out/bench_native_tc.err
timeout 600 python bench.py --steps 3 --warmup 2 --engine native_tf32 --no-cpu-baseline > gpurun_out/bench_native_tf32.json 2> gpurun_out/bench_native_tf32.err
cat gpurun_out/diag_d.log | tail -8 | cut -c1-400; grep -E "passed|failed" gpurun_out/tests_d_conv.log gpurun_out/tests_d_rest.log | tail -3; head -c 300 gpurun_out/bench_native_tc.json; tail -2 gpurun_out/bench_native_tc.err
