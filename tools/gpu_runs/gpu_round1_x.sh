#!/bin/bash
# split groups sweep, pooled max-pool mask, slots rule: conv tests, bench sweep, attack tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x > gpurun_out/tests_x_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_x_conv.log
tail -4 gpurun_out/tests_x_conv.log | cut -c1-300
for G in 2 1 4; do I2V_TC_SPLIT_GROUPS=$G timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_x_g$G.json 2> gpurun_out/bench_x_g$G.err; done
I2V_TC_SPLIT_GROUPS=4 timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x -k "conv_tc" > gpurun_out/tests_x_conv_g4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_x_conv_g4.log
tail -3 gpurun_out/tests_x_conv_g4.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_attacks.py -m gpu -q --timeout 600 -x > gpurun_out/tests_x_att.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_x_att.log
tail -4 gpurun_out/tests_x_att.log | cut -c1-300
python - <<'PY'
import json
for f in ('bench_x_g2','bench_x_g1','bench_x_g4'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:7]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- G=2; head -22 gpurun_out/bench_x_g2.err
echo ---- G=4; head -12 gpurun_out/bench_x_g4.err
