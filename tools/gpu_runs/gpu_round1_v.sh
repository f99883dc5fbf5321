#!/bin/bash
# full gpu suite + smoke + default bench (both arms) + ncu launch list and full capture at the bench's launch size
mkdir -p gpurun_out /tmp/prof
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/tests_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_v.log
tail -6 gpurun_out/tests_v.log | cut -c1-300
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_v.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_v.log; tail -3 gpurun_out/smoke_v.log
( time timeout 600 python bench.py --shapes ) > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_v_ref.json 2> gpurun_out/bench_v_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v.csv python bench.py --steps 1 --warmup 1 --clips 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_v_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_persist|stem_|cosine|adam_compose|maxpool|std_" -s 60 -c 70 -o /tmp/prof/full python bench.py --steps 1 --warmup 1 --clips 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_v_full.log 2>&1
ncu -i /tmp/prof/full.ncu-rep --page raw --csv > gpurun_out/full_raw_v.csv 2>/dev/null
ls -la /tmp/prof | tail -3
python - <<'PY'
import json
for f in ('bench_v','bench_v_ref'):
    try:
        r=[json.loads(l) for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0]
        print(f, r['n_gpus'], round(r['value'],1), round(r['ms_per_step'],1), 'e2e', r['e2e'] and round(r['e2e']['value'],1), r.get('clocks'), r.get('cpu_baseline'))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-1200:])
PY
head -24 gpurun_out/bench_v.err
