#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/tests_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_b.log
timeout 300 python tools/microbench.py --out gpurun_out/microbench_b.json --clusters default,8,16 > gpurun_out/microbench_b.log 2>&1
I2V_COS_CTAS_PER_SM=1 timeout 300 python tools/microbench.py --clusters 16,8 > gpurun_out/microbench_b_1cta.log 2>&1
tail -15 gpurun_out/tests_b.log
