#!/bin/bash
# templated max-pool kernels, restructured im2col, parallel K1 reductions: full GPU suite, bench + shapes, dual-issuer threshold A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/tests_ac.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_ac.log
tail -6 gpurun_out/tests_ac.log | cut -c1-400
for MK in 8 4 2; do I2V_TC_ALO_MINKIT=$MK timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_ac_mk$MK.json 2> gpurun_out/bench_ac_mk$MK.err; done
python - <<'PY'
import json
for f in ('bench_ac_mk8','bench_ac_mk4','bench_ac_mk2'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step']): print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- mk8; head -24 gpurun_out/bench_ac_mk8.err
echo ---- mk4; head -24 gpurun_out/bench_ac_mk4.err
echo ---- mk2; head -24 gpurun_out/bench_ac_mk2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_ac.csv python bench.py --clips 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ac_launch.log 2>&1; echo "ncu launches rc=$?"
