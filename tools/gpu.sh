#!/bin/bash
# build in-tree (the .so travels with the snapshot), then run a recipe from tools/gpu_runs/ on a GPU box:
#   tools/gpu.sh [--gpus N] <timeout-seconds> <script>
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; exit 9; }
exec /usr/local/graft/bin/gpurun $GP --timeout "$1" -- "bash $2"
