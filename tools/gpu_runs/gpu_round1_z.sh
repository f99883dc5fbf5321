#!/bin/bash
# re-entry check of HEAD: full GPU test suite, bench (default line) + per-shape table
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x --durations=15 > gpurun_out/tests_z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_z.log
tail -25 gpurun_out/tests_z.log | cut -c1-300
timeout 500 python bench.py --steps 10 --warmup 3 --shapes > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench_z.json')); print(round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'], 'e2e', round(r['e2e']['value']), r['roofline']['frac'])
    for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step']): print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
except Exception as e: print('ERR',e, open('gpurun_out/bench_z.err').read()[-1500:])
PY
head -30 gpurun_out/bench_z.err
