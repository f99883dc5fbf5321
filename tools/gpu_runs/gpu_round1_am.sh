#!/bin/bash
# dual MMA issuers with the register epilogue (strided data-gradient classes): conv tests, bench A/B with shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 > gpurun_out/tests_am.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_am.log
tail -5 gpurun_out/tests_am.log | cut -c1-400
for A in 1 0; do I2V_TC_ALO_REGEPI=$A timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_am_a$A.json 2> gpurun_out/bench_am_a$A.err; done
python - <<'PY'
import json
for f in ('bench_am_a1','bench_am_a0'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); v=r['roofline_all']['i2v_conv_tc_dgrad_class_f32']
        print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'], 'classes avg %.1f us share %.3f'%(v['avg_us'], v['share_of_step']), r['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
