#!/bin/bash
# dual MMA issuers (A_lo in tensor memory, cross2 accumulator): conv tests, bench shapes, control
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x > gpurun_out/tests_aa_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_aa_conv.log
tail -12 gpurun_out/tests_aa_conv.log | cut -c1-300
for A in 1 0; do I2V_TC_ALO_TMEM=$A timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_aa_a$A.json 2> gpurun_out/bench_aa_a$A.err; done
timeout 900 python -m pytest tests/test_gpu_attacks.py -m gpu -q --timeout 600 -x > gpurun_out/tests_aa_att.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_aa_att.log
tail -4 gpurun_out/tests_aa_att.log | cut -c1-300
python - <<'PY'
import json
for f in ('bench_aa_a1','bench_aa_a0'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:7]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-1500:])
PY
echo ---- dual; head -22 gpurun_out/bench_aa_a1.err
