#!/bin/bash
# DEEP = both A operands of the BN = 64 dual-issuer kernel in tensor memory: conv tests, bench A/B with shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 300 2>&1 | tail -3 | cut -c1-300
for V in 1 0; do I2V_TC_DEEP=$V timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_au_$V.json 2> gpurun_out/bench_au_$V.err; done
python - <<'PY'
import json
for f in ('bench_au_1','bench_au_0'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'], r['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-600:])
PY
for f in 1 0; do echo "--- deep=$f"; grep -E "64->64 k3s1" gpurun_out/bench_au_$f.err; done
