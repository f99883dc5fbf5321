#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_trace.py --frames 32 --layers "l1.conv1(64->64,l1.conv3(64->256,1x1),l1.conv3(64->256,1x1)+res,l1.conv2,l2.conv1,l2.conv2" > gpurun_out/tc_trace_m.log 2>&1
timeout 300 python tools/tc_probe.py --frames 128 > gpurun_out/tc_probe_m128.log 2>&1
for C in 64 128; do I2V_CHUNK=$C timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_m_chunk$C.json 2> gpurun_out/bench_m_chunk$C.err; done
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_attacks.py -m gpu -q --timeout 600 -x > gpurun_out/tests_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_m.log
tail -5 gpurun_out/tests_m.log
cat gpurun_out/tc_trace_m.log | head -150
python - <<'PY'
import json
for l in open('gpurun_out/tc_probe_m128.log'):
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-30s K=%-5d x3 %.3f ms %6.1f TF %5.0f GB/s | x1 %.3f ms | cudnn tf32 %.3f'%(r['layer'],r['K'],r['ms_tc_x3'],r['tflops_tc_x3'],r['gbs_tc_x3'],r['ms_tc_x1'],r['ms_cudnn_tf32']))
for f in ('bench_m_chunk64','bench_m_chunk128'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-800:])
PY
