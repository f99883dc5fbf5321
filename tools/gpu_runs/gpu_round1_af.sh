#!/bin/bash
# third MMA issuer for the K-heavy BN = 64 im2col layers: full GPU suite, bench A/B with shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/tests_af.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_af.log
tail -8 gpurun_out/tests_af.log | cut -c1-600
for TRI in 1 0; do I2V_TC_TRI=$TRI timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_af_tri$TRI.json 2> gpurun_out/bench_af_tri$TRI.err; done
python - <<'PY'
import json
for f in ('bench_af_tri1','bench_af_tri0'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- tri1; head -8 gpurun_out/bench_af_tri1.err
echo ---- tri0; head -8 gpurun_out/bench_af_tri0.err
