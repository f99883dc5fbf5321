#!/bin/bash
# re-entry baseline: full gpu test suite, smoke, default bench (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/n_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x ) > gpurun_out/tests_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_n.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_n.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_n.log
( time timeout 600 python bench.py ) > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_n_ref.json 2> gpurun_out/bench_n_ref.err
tail -4 gpurun_out/tests_n.log; tail -3 gpurun_out/smoke_n.log
cat gpurun_out/bench_n.json | cut -c1-3000; tail -5 gpurun_out/bench_n.err
cat gpurun_out/bench_n_ref.json | cut -c1-1000; tail -5 gpurun_out/bench_n_ref.err
