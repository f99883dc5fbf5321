#!/bin/bash
# DEEP pipeline layout (BN = 64 dual issuer, 1 accumulator stage, 8-slot A_lo ring): conv tests, bench A/B, with the direct stem
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 2>&1 | tail -3 | cut -c1-300
for V in "1 0" "0 0" "1 1"; do set -- $V; I2V_TC_DEEP=$1 I2V_STEM_DIRECT=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_at_$1$2.json 2> gpurun_out/bench_at_$1$2.err; done
python - <<'PY'
import json
for f in ('bench_at_10','bench_at_00','bench_at_11'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); v=r['roofline_all']['i2v_conv_stem_fwd_f32']
        print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'], 'stem fwd avg %.1f us'%v['avg_us'], r['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-600:])
PY
for f in 10 00; do echo "--- deep,direct=$f"; grep -E "64->64 k3s1" gpurun_out/bench_at_$f.err; done
