#!/bin/bash
# TemporalTranslation / TAP / ILAF (K8, K9, K3d): full GPU suite + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/tests_ad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_ad.log
tail -30 gpurun_out/tests_ad.log | cut -c1-600
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench_ad.json')); print(round(r['value']), round(r['ms_per_step'],2), 'e2e', round(r['e2e']['value']), r['roofline']['frac'], r['cpu_baseline'])
except Exception as e: print('ERR',e, open('gpurun_out/bench_ad.err').read()[-1500:])
PY
