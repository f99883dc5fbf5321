#!/bin/bash
# round-1 final evidence: full GPU suite, smoke, both bench arms, ncu launch list + --set full capture at the bench's launch
# size (256-frame chunk), pipeline trace of the tensor-core kernel
mkdir -p gpurun_out /tmp/prof
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/tests_ag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_ag.log
tail -6 gpurun_out/tests_ag.log | cut -c1-300
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_ag.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_ag.log; tail -3 gpurun_out/smoke_ag.log
( time timeout 600 python bench.py --shapes ) > gpurun_out/bench_ag.json 2> gpurun_out/bench_ag.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_ag_ref.json 2> gpurun_out/bench_ag_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_ag.csv python bench.py --steps 1 --warmup 1 --clips 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ag_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_persist|stem_|cosine|adam_compose|maxpool" -s 64 -c 72 -o /tmp/prof/full python bench.py --steps 1 --warmup 1 --clips 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ag_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof/full.ncu-rep --page raw --csv > gpurun_out/full_raw_ag.csv 2>/dev/null
ls -la /tmp/prof | tail -2
timeout 300 python tools/tc_trace.py --frames 256 --tiles 8 --layers "l1.conv2,l1.conv3(64->256,1x1)+res,l1.conv1(256,l2.conv2,l2.conv3(128->512,1x1)+res" > gpurun_out/tc_trace_ag.txt 2>&1; echo "trace rc=$?"
python - <<'PY'
import json
for f in ('bench_ag','bench_ag_ref'):
    try:
        r=[json.loads(l) for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0]
        print(f, r['n_gpus'], round(r['value'],1), round(r['ms_per_step'],1), 'e2e', r['e2e'] and round(r['e2e']['value'],1), r.get('clocks'), r.get('cpu_baseline'))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-1200:])
PY
