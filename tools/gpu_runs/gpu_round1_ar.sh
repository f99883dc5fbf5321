#!/bin/bash
# EXPERIMENTAL direct first-layer forward: kernel test, then bench A/B behind $I2V_STEM_DIRECT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 200 -x -k "stem_fwd_direct" 2>&1 | tail -25 | cut -c1-300
for D in 1 0; do I2V_STEM_DIRECT=$D timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ar_d$D.json 2> gpurun_out/bench_ar_d$D.err; done
python - <<'PY'
import json
for f in ('bench_ar_d1','bench_ar_d0'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); v=r['roofline_all']['i2v_conv_stem_fwd_f32']
        print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'], 'stem fwd avg %.1f us'%v['avg_us'], r['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-600:])
PY
