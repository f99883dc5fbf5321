#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 > gpurun_out/tests_h_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_h_conv.log
timeout 600 python tools/tc_probe.py --frames 32 --out gpurun_out/tc_probe_h.json > gpurun_out/tc_probe_h.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --engine native --no-cpu-baseline > gpurun_out/bench_h_native.json 2> gpurun_out/bench_h_native.err
timeout 600 python bench.py --steps 3 --warmup 2 --engine native_tf32 --no-cpu-baseline > gpurun_out/bench_h_native_tf32.json 2> gpurun_out/bench_h_native_tf32.err
grep -E "passed|failed" gpurun_out/tests_h_conv.log | tail -2; grep -E "^(FAILED|E  )" gpurun_out/tests_h_conv.log | head -20 | cut -c1-220
python - <<'PY'
import json
for l in open('gpurun_out/tc_probe_h.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-28s K=%-5d x3 %.3f ms %6.1f TF %5.0f GB/s err %.1e | x1 %.3f ms %6.1f TF | simt %.3f | cudnn fp32 %.3f tf32 %.3f'%(r['layer'],r['K'],r['ms_tc_x3'],r['tflops_tc_x3'],r['gbs_tc_x3'],r['err_max_x3'],r['ms_tc_x1'],r['tflops_tc_x1'],r['ms_simt'],r['ms_cudnn_fp32'],r['ms_cudnn_tf32']))
for f in ('bench_h_native','bench_h_native_tf32'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), round(r['e2e']['value']), r['config']['final_cost'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
