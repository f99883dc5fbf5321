#!/bin/bash
# variable epilogue slots (third pipeline stage for K-heavy layers): conv tests, bench with shapes, slots=2 control
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x > gpurun_out/tests_w_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_w_conv.log
tail -4 gpurun_out/tests_w_conv.log | cut -c1-300
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err
I2V_TC_EPI_SLOTS=2 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_w_s2.json 2> gpurun_out/bench_w_s2.err
I2V_TC_EPI_SLOTS=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_w_s1.json 2> gpurun_out/bench_w_s1.err
timeout 900 python -m pytest tests/test_gpu_attacks.py -m gpu -q --timeout 600 -x > gpurun_out/tests_w_att.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_w_att.log
tail -4 gpurun_out/tests_w_att.log | cut -c1-300
python - <<'PY'
import json
for f in ('bench_w','bench_w_s2','bench_w_s1'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- default; head -22 gpurun_out/bench_w.err
echo ---- slots=2; head -22 gpurun_out/bench_w_s2.err
echo ---- slots=1; head -22 gpurun_out/bench_w_s1.err
