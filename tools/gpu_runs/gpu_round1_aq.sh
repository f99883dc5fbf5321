#!/bin/bash
# warp-parallel partial combines (K6, K9): kernel + DR / ILAF attack tests, isolated timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "std_loss or ila or dispersion or video_variants" 2>&1 | tail -3
timeout 600 python tools/microbench.py --frames 256 --iters 8 --clusters default --out gpurun_out/microbench_aq.json 2>&1 | grep -E "ila|std_acc" | cut -c1-200
