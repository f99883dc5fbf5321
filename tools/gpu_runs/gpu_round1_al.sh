#!/bin/bash
# N = 2: bench line under torchrun (weak scaling, no collective) + the multi-GPU checks (clip sharding bit-identical, ensemble placement)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_al_n2.json 2> gpurun_out/bench_al_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    r=[json.loads(l) for l in open('gpurun_out/bench_al_n2.json') if l.startswith('{')][0]
    print(r['n_gpus'], round(r['value']), round(r['ms_per_step'],2), 'e2e', round(r['e2e']['value']), r['clocks'])
except Exception as e: print('ERR',e, open('gpurun_out/bench_al_n2.err').read()[-1500:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/multigpu_check.py > gpurun_out/multigpu_al.log 2>&1; echo "multigpu rc=$?"; tail -5 gpurun_out/multigpu_al.log | cut -c1-600
