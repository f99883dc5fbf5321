#!/bin/bash
# chunk-size sweep (L2 residency), cuDNN comparison, per-layer probe, ncu launch list + full capture for traffic
mkdir -p gpurun_out /tmp/prof
timeout 300 python tools/tc_probe.py --frames 32 --out gpurun_out/tc_probe_o32.json > gpurun_out/tc_probe_o32.log 2>&1
timeout 300 python tools/tc_probe.py --frames 12 --out gpurun_out/tc_probe_o12.json > gpurun_out/tc_probe_o12.log 2>&1
for C in 8 12 16 24; do I2V_CHUNK=$C timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_o_chunk$C.json 2> gpurun_out/bench_o_chunk$C.err; done
for E in cudnn cudnn_tf32 native_tf32; do timeout 300 python bench.py --steps 5 --warmup 3 --engine $E --no-cpu-baseline --no-e2e > gpurun_out/bench_o_$E.json 2> gpurun_out/bench_o_$E.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_o.csv python bench.py --steps 2 --warmup 1 --clips 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_o_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_persist|stem_|cosine|adam_compose|maxpool" -s 0 -c 60 -o /tmp/prof/full python bench.py --steps 1 --warmup 1 --clips 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_o_full.log 2>&1
ncu -i /tmp/prof/full.ncu-rep --page raw --csv > gpurun_out/full_raw_o.csv 2>/dev/null
ls -la /tmp/prof gpurun_out | tail -30
python - <<'PY'
import json
for f in ('tc_probe_o32','tc_probe_o12'):
    print(f)
    for l in open('gpurun_out/%s.log'%f):
        try: r=json.loads(l)
        except Exception: print(l.strip()[:200]); continue
        print('%-30s K=%-5d x3 %.3f ms %6.1f TF %5.0f GB/s | x1 %.3f ms | simt %.3f | cudnn fp32 %.3f tf32 %.3f'%(r['layer'],r['K'],r['ms_tc_x3'],r['tflops_tc_x3'],r['gbs_tc_x3'],r['ms_tc_x1'],r['ms_simt'],r['ms_cudnn_fp32'],r['ms_cudnn_tf32']))
for f in ('chunk8','chunk12','chunk16','chunk24','cudnn','cudnn_tf32','native_tf32'):
    try:
        r=json.load(open('gpurun_out/bench_o_%s.json'%f)); print(f, round(r['value']), round(r['ms_per_step'],1))
        for k,v in sorted(r['roofline_all'].items(), key=lambda kv:-kv[1]['share_of_step'])[:6]: print('   %-32s share %.3f n=%d avg %.1f us  %.0f GB/s'%(k,v['share_of_step'],v['launches'],v['avg_us'],v['achieved']))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/bench_o_%s.err'%f).read()[-800:])
PY
