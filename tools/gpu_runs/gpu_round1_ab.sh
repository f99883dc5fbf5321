#!/bin/bash
# per-row max-pool kernels (mark_dead), unrolled col2im, TF32-off white-box gradients: full GPU suite, bench + shapes,
# launch list and --set full captures of the memory-bound helper kernels at one 256-frame chunk
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > gpurun_out/tests_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_ab.log
tail -16 gpurun_out/tests_ab.log | cut -c1-400
timeout 500 python bench.py --steps 10 --warmup 3 --shapes > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench_ab.json')); print(round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'], 'e2e', round(r['e2e']['value']), r['roofline']['frac'])
    for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step']): print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
except Exception as e: print('ERR',e, open('gpurun_out/bench_ab.err').read()[-1500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_ab.csv python bench.py --clips 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ab_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'im2col|col2im|maxpool|cosine_loss' -c 10 -o gpurun_out/helpers_ab -f python bench.py --clips 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ab_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/ | tail -8
