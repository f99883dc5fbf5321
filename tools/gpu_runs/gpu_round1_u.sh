#!/bin/bash
# 2 GPUs: multi-GPU checks (sharding, one backbone per GPU over NCCL), bench at N=2, then N=1 both arms
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/u_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/multigpu_u.json 2> gpurun_out/multigpu_u.err; echo "multigpu rc=$?"
cat gpurun_out/multigpu_u.json; tail -5 gpurun_out/multigpu_u.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_u_n2.json 2> gpurun_out/bench_u_n2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_u_n1.json 2> gpurun_out/bench_u_n1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for f in ('bench_u_n2','bench_u_n1'):
    try:
        r=json.load(open('gpurun_out/%s.json'%f)); print(f, r['n_gpus'], round(r['value']), round(r['ms_per_step'],1), 'e2e', r['e2e'] and round(r['e2e']['value']), r['config']['chunk_frames'], r['clocks'], r.get('cpu_baseline'))
        print('   roofline', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r['roofline'].items() if k in ('kernel','achieved','peak','frac','traffic','avg_us','share_of_step')})
    except Exception as e: print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-1200:])
PY
