#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 > gpurun_out/tests_e_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_e_conv.log
timeout 600 python tools/tc_probe.py --frames 32 --out gpurun_out/tc_probe_e.json > gpurun_out/tc_probe_e.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "zero_features" > gpurun_out/tests_e_k.log 2>&1
grep -E "passed|failed" gpurun_out/tests_e_conv.log | tail -2; grep -E "^(FAILED|E  )" gpurun_out/tests_e_conv.log | head -20 | cut -c1-200; cut -c1-700 gpurun_out/tc_probe_e.log | tail -12
