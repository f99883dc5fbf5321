#!/bin/bash
# session-2 first pass: is the persistent tcgen05 conv kernel correct, what does the native engine deliver
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 600 > gpurun_out/tests_i_conv.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-T0 ))" >> gpurun_out/tests_i_conv.log
timeout 400 python bench.py --steps 3 --warmup 3 --engine native --no-cpu-baseline > gpurun_out/bench_i_native.json 2> gpurun_out/bench_i_native.err
timeout 300 python bench.py --steps 3 --warmup 3 --engine native_tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_i_native_tf32.json 2> gpurun_out/bench_i_native_tf32.err
timeout 300 python bench.py --steps 3 --warmup 3 --engine cudnn --no-cpu-baseline --no-e2e > gpurun_out/bench_i_cudnn.json 2> gpurun_out/bench_i_cudnn.err
timeout 300 python bench.py --steps 3 --warmup 3 --engine cudnn_tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_i_cudnn_tf32.json 2> gpurun_out/bench_i_cudnn_tf32.err
timeout 400 python tools/tc_probe.py --frames 32 --out gpurun_out/tc_probe_i.json > gpurun_out/tc_probe_i.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_i_native.csv python bench.py --steps 1 --warmup 1 --clips 2 --engine native --no-cpu-baseline --no-e2e > gpurun_out/ncu_i.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_attacks.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/tests_i_rest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-T0 ))" >> gpurun_out/tests_i_rest.log
grep -E "passed|failed|rc=" gpurun_out/tests_i_conv.log gpurun_out/tests_i_rest.log | tail -6
grep -E "^(FAILED|E  )" gpurun_out/tests_i_conv.log gpurun_out/tests_i_rest.log | head -30 | cut -c1-220
python - <<'PY'
import json
for l in open('gpurun_out/tc_probe_i.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-28s K=%-5d x3 %.3f ms %6.1f TF %5.0f GB/s err %.1e | x1 %.3f ms %6.1f TF | simt %.3f | cudnn fp32 %.3f tf32 %.3f'%(r['layer'],r['K'],r['ms_tc_x3'],r['tflops_tc_x3'],r['gbs_tc_x3'],r['err_max_x3'],r['ms_tc_x1'],r['tflops_tc_x1'],r['ms_simt'],r['ms_cudnn_fp32'],r['ms_cudnn_tf32']))
for f in ('bench_i_native','bench_i_native_tf32','bench_i_cudnn','bench_i_cudnn_tf32'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1), r['e2e'] and round(r['e2e']['value']), r['config']['final_cost'], r['clocks'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
