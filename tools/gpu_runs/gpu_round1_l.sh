#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 -s > gpurun_out/tests_l_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_l_conv.log
timeout 400 python tools/tc_probe.py --frames 32 --out gpurun_out/tc_probe_l.json > gpurun_out/tc_probe_l.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l_native.json 2> gpurun_out/bench_l_native.err
timeout 1200 python -m pytest tests/test_gpu_attacks.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/tests_l_rest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_l_rest.log
grep -E "passed|failed|rc=" gpurun_out/tests_l_conv.log gpurun_out/tests_l_rest.log | tail -6
grep -E "^(FAILED|E  )" gpurun_out/tests_l_conv.log gpurun_out/tests_l_rest.log | head -30 | cut -c1-250
grep "engine parity" gpurun_out/tests_l_conv.log | cut -c1-200
python - <<'PY'
import json
for l in open('gpurun_out/tc_probe_l.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-30s K=%-5d x3 %.3f ms %6.1f TF %5.0f GB/s err %.1e | x1 %.3f ms %6.1f TF | simt %.3f | cudnn fp32 %.3f tf32 %.3f'%(r['layer'],r['K'],r['ms_tc_x3'],r['tflops_tc_x3'],r['gbs_tc_x3'],r['err_max_x3'],r['ms_tc_x1'],r['tflops_tc_x1'],r['ms_simt'],r['ms_cudnn_fp32'],r['ms_cudnn_tf32']))
for f in ('bench_l_native',):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['e2e'] and round(r['e2e']['value']), r['config']['final_cost'], r['clocks'])
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step']): print('   %-32s share %.3f avg %.1f us  %.0f GB/s (%.2f)  %s'%(k,v['share_of_step'],v['avg_us'],v['achieved'],v['frac'], ('%.0f TF alg'%v['tensor']['achieved_algorithmic']) if 'tensor' in v else ''))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-1500:])
PY
