#!/bin/bash
# first GPU pass: parity tests, isolated kernel timing, bench line, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
timeout 300 python tools/microbench.py --out gpurun_out/microbench.json > gpurun_out/microbench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cudnn.json 2> gpurun_out/bench_cudnn.err
timeout 300 python bench.py --steps 5 --warmup 3 --engine cudnn_tf32 --no-cpu-baseline > gpurun_out/bench_cudnn_tf32.json 2> gpurun_out/bench_cudnn_tf32.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --clips 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"cosine_loss_grad|adam_compose" -c 4 -o gpurun_out/prof_r1_membound python bench.py --steps 2 --warmup 1 --clips 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/tests.log; cat gpurun_out/bench_cudnn.json | head -c 1500
