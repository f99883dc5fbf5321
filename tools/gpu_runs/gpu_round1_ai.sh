#!/bin/bash
# K1 cluster-size sweep with the L2 hints (isolated, 256 frames)
for C in default 8 4; do echo "--- cluster=$C"; timeout 200 python tools/microbench.py --frames 256 --iters 8 --clusters $C 2>&1 | grep "cosine_loss_grad_f32\"" | cut -c1-330; done
echo "--- 1 CTA per SM"; I2V_COS_CTAS_PER_SM=1 timeout 200 python tools/microbench.py --frames 256 --iters 8 --clusters default 2>&1 | grep "cosine_loss_grad_f32\"" | cut -c1-330
