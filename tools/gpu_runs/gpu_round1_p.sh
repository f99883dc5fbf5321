#!/bin/bash
# TMA epilogue (v3) + bit masks: conv tests, per-layer probe, trace, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -x > gpurun_out/tests_p_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_p_conv.log
tail -15 gpurun_out/tests_p_conv.log | cut -c1-300
timeout 300 python tools/tc_probe.py --frames 32 --out gpurun_out/tc_probe_p32.json > gpurun_out/tc_probe_p32.log 2>&1
timeout 300 python tools/tc_trace.py --frames 32 --layers "l1.conv3(64->256,1x1)+res,l2.conv3(128->512,1x1)+res,l2.conv2" --tiles 8 > gpurun_out/tc_trace_p.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err
I2V_NATIVE_BITS=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_p_nobits.json 2> gpurun_out/bench_p_nobits.err
timeout 900 python -m pytest tests/test_gpu_attacks.py -m gpu -q --timeout 600 -x > gpurun_out/tests_p_att.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_p_att.log
tail -5 gpurun_out/tests_p_att.log | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/tc_probe_p32.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-30s K=%-5d x3 %.3f ms %6.1f TF %5.0f GB/s err %.1e | x1 %.3f ms | cudnn tf32 %.3f'%(r['layer'],r['K'],r['ms_tc_x3'],r['tflops_tc_x3'],r['gbs_tc_x3'],r['err_max_x3'],r['ms_tc_x1'],r['ms_cudnn_tf32']))
for f in ('bench_p','bench_p_nobits'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['config']['final_cost'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:6]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
head -60 gpurun_out/tc_trace_p.log
