#!/bin/bash
# configs 3-5 through the public classes; first-layer dgrad GEMM with 128-row padding (BN = 128 dual issuer) A/B
mkdir -p gpurun_out
timeout 600 python tools/configs_bench.py --out gpurun_out/configs_n1.json > gpurun_out/configs_n1.log 2> gpurun_out/configs_n1.err; echo "configs rc=$?"
cat gpurun_out/configs_n1.log | cut -c1-400; tail -5 gpurun_out/configs_n1.err
for NZ in 64 128; do I2V_STEM_NZ=$NZ timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ae_nz$NZ.json 2> gpurun_out/bench_ae_nz$NZ.err; done
python - <<'PY'
import json
for f in ('bench_ae_nz64','bench_ae_nz128'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:4]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
I2V_STEM_NZ=128 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -k "stem" 2>&1 | tail -3
