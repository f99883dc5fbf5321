#!/bin/bash
# final state check: full GPU suite, smoke, default bench line with the per-shape table
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/tests_an.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_an.log
tail -6 gpurun_out/tests_an.log | cut -c1-300
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_an.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_an.log; tail -2 gpurun_out/smoke_an.log
( timeout 600 python bench.py --shapes ) > gpurun_out/bench_an.json 2> gpurun_out/bench_an.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    r=[json.loads(l) for l in open('gpurun_out/bench_an.json') if l.startswith('{')][0]
    print(r['n_gpus'], round(r['value'],1), round(r['ms_per_step'],2), 'e2e', round(r['e2e']['value'],1), r['clocks'], 'roofline', r['roofline']['frac'], r['roofline']['traffic'])
    for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step']): print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s frac %.3f'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved'],v['frac']))
except Exception as e: print('ERR',e, open('gpurun_out/bench_an.err').read()[-1200:])
PY
