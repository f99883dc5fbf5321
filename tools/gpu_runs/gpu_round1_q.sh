#!/bin/bash
# L2 prefetch + stem dgrad on tensor cores: conv tests, in-situ per-shape table, prefetch sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x > gpurun_out/tests_q_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_q_conv.log
tail -15 gpurun_out/tests_q_conv.log | cut -c1-300
for PF in 0 2 4; do I2V_TC_PREFETCH=$PF timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --shapes > gpurun_out/bench_q_pf$PF.json 2> gpurun_out/bench_q_pf$PF.err; done
I2V_NATIVE_STEM_TC=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_q_nostemtc.json 2> gpurun_out/bench_q_nostemtc.err
timeout 300 python tools/tc_probe.py --frames 32 --out gpurun_out/tc_probe_q32.json > gpurun_out/tc_probe_q32.log 2>&1
timeout 900 python -m pytest tests/test_gpu_attacks.py -m gpu -q --timeout 600 -x > gpurun_out/tests_q_att.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_q_att.log
tail -5 gpurun_out/tests_q_att.log | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/tc_probe_q32.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-30s K=%-5d x3 %.3f ms %6.1f TF %5.0f GB/s err %.1e | x1 %.3f ms | cudnn tf32 %.3f'%(r['layer'],r['K'],r['ms_tc_x3'],r['tflops_tc_x3'],r['gbs_tc_x3'],r['err_max_x3'],r['ms_tc_x1'],r['ms_cudnn_tf32']))
for f in ('bench_q_pf0','bench_q_pf2','bench_q_pf4','bench_q_nostemtc'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:6]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
echo ---- pf2 shapes; cat gpurun_out/bench_q_pf2.err | head -40
echo ---- pf0 shapes; cat gpurun_out/bench_q_pf0.err | head -40
