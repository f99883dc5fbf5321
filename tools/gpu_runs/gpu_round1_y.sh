#!/bin/bash
# A_lo in tensor memory (TS MMA): conv tests, bench shapes, control with ALO off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x -k "conv_tc or stem or engine" > gpurun_out/tests_y_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_y_conv.log
tail -12 gpurun_out/tests_y_conv.log | cut -c1-300
for A in 1 0 2; do I2V_TC_ALO_TMEM=$A timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_y_a$A.json 2> gpurun_out/bench_y_a$A.err; done
python - <<'PY'
import json
for f in ('bench_y_a1','bench_y_a0','bench_y_a2'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- ALO=1; head -22 gpurun_out/bench_y_a1.err
echo ---- ALO=0; head -22 gpurun_out/bench_y_a0.err
echo ---- ALO=2; head -22 gpurun_out/bench_y_a2.err
