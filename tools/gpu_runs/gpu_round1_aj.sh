#!/bin/bash
# K1 with the split cluster barrier and 8-CTA clusters: kernel tests, isolated timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "cosine" 2>&1 | tail -3
timeout 200 python tools/microbench.py --frames 256 --iters 8 --clusters default,16 2>&1 | grep "cosine_loss_grad" | cut -c1-330
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_aj.json 2> gpurun_out/bench_aj.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/bench_aj.json')); v=r['roofline_all']['i2v_cosine_loss_grad_f32']
print(round(r['value']), round(r['ms_per_step'],2), 'K1 avg %.1f us  %.0f GB/s  frac %.3f'%(v['avg_us'],v['achieved'],v['frac']), r['clocks'])
PY
