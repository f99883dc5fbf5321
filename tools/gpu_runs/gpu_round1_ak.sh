#!/bin/bash
# 512-frame chunks (one chunk per GPU batch) against the 256-frame default
mkdir -p gpurun_out
for CH in 512 256; do I2V_CHUNK=$CH I2V_STEM_GROUP_MB=8192 timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ak_c$CH.json 2> gpurun_out/bench_ak_c$CH.err; done
python - <<'PY'
import json
for f in ('bench_ak_c512','bench_ak_c256'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],2), r['config']['final_cost'], r['config'].get('chunk_frames'), r['clocks'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:5]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-1500:])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
