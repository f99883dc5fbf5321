#!/bin/bash
# K1 with L2 eviction hints: kernel tests, A/B in the bench (hints on / off), isolated microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "cosine" 2>&1 | tail -3
for H in 1 0; do I2V_COS_L2_HINTS=$H timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ah_h$H.json 2> gpurun_out/bench_ah_h$H.err; done
python - <<'PY'
import json
for f in ('bench_ah_h1','bench_ah_h0'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); v=r['roofline_all']['i2v_cosine_loss_grad_f32']
        print(f, round(r['value']), round(r['ms_per_step'],2), 'K1 avg %.1f us  %.0f GB/s  frac %.3f'%(v['avg_us'],v['achieved'],v['frac']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
for H in 1 0; do echo "--- microbench hints=$H"; I2V_COS_L2_HINTS=$H timeout 300 python tools/microbench.py --frames 256 --iters 8 --clusters default 2>&1 | grep -i "cosine\|K1" | cut -c1-400; done
