#!/bin/bash
# isolated timing of every memory-bound kernel (K1, K3a-d, K6, K8, K9, max pooling)
mkdir -p gpurun_out
timeout 600 python tools/microbench.py --frames 256 --iters 8 --clusters default --out gpurun_out/microbench_ap.json 2>&1 | cut -c1-260
