#!/bin/bash
# col2im with affine tap addressing: stem tests + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -k "stem or engine" 2>&1 | tail -3
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ao.json 2> gpurun_out/bench_ao.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/bench_ao.json')); v=r['roofline_all']['i2v_conv_stem_dgrad_f32']
print(round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'], 'stem dgrad avg %.1f us'%v['avg_us'], r['clocks'])
PY
