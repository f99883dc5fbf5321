#!/bin/bash
# B_lo on the fly + stem fwd TC + frame-grouped stems; chunk sweep for wave quantisation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x > gpurun_out/tests_r_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_r_conv.log
tail -15 gpurun_out/tests_r_conv.log | cut -c1-300
for C in 32 24 48 96; do I2V_CHUNK=$C timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_r_c$C.json 2> gpurun_out/bench_r_c$C.err; done
timeout 900 python -m pytest tests/test_gpu_attacks.py -m gpu -q --timeout 600 -x > gpurun_out/tests_r_att.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_r_att.log
tail -5 gpurun_out/tests_r_att.log | cut -c1-300
python - <<'PY'
import json
for f in ('bench_r_c32','bench_r_c24','bench_r_c48','bench_r_c96'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:7]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- c32 shapes; cat gpurun_out/bench_r_c32.err | head -24
echo ---- c48 shapes; cat gpurun_out/bench_r_c48.err | head -24
