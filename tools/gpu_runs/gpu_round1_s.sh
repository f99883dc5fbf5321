#!/bin/bash
# full gpu suite (DR attack, forced max-pool winners, reverted B_lo), big-chunk sweep
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/tests_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_s.log
tail -25 gpurun_out/tests_s.log | cut -c1-300
for C in 64 128 256; do I2V_CHUNK=$C timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_s_c$C.json 2> gpurun_out/bench_s_c$C.err; done
python - <<'PY'
import json
for f in ('bench_s_c64','bench_s_c128','bench_s_c256'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:7]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- c128 shapes; cat gpurun_out/bench_s_c128.err | head -24
