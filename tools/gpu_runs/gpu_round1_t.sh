#!/bin/bash
# DR fixes + memory-sized chunks: failing tests again, default bench, stem group sweep
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_attacks.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 ) > gpurun_out/tests_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_t.log
tail -8 gpurun_out/tests_t.log | cut -c1-300
for G in 48 100000; do I2V_STEM_GROUP_MB=$G timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_t_g$G.json 2> gpurun_out/bench_t_g$G.err; done
I2V_CHUNK=512 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_t_c512.json 2> gpurun_out/bench_t_c512.err
python - <<'PY'
import json
for f in ('bench_t_g48','bench_t_g100000','bench_t_c512'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['chunk_frames'], r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:7]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
