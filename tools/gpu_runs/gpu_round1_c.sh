#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/tests_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_c.log
timeout 300 python tools/microbench.py --out gpurun_out/microbench_c.json --clusters default,8 > gpurun_out/microbench_c.log 2>&1
I2V_COS_NONPERSISTENT=1 timeout 300 python tools/microbench.py --clusters default > gpurun_out/microbench_c_nonpersist.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --engine native --no-cpu-baseline > gpurun_out/bench_native_simt.json 2> gpurun_out/bench_native_simt.err
grep -E "passed|failed" gpurun_out/tests_c.log | tail -3; grep cosine gpurun_out/microbench_c.log | head -3; head -c 600 gpurun_out/bench_native_simt.json; tail -3 gpurun_out/bench_native_simt.err
