#!/bin/bash
mkdir -p gpurun_out
for P in 1 0; do
  echo "=== PERSISTENT=$P vgg 3 32" ; I2V_TC_PERSISTENT=$P timeout 300 python tools/diag_race.py vgg 3 32 3 12 2>&1 | grep -E "<<<<|gimg|SUSPECT|Error|error" | head -40
  echo "=== PERSISTENT=$P resnet 2 64" ; I2V_TC_PERSISTENT=$P timeout 300 python tools/diag_race.py resnet 2 64 2 12 2>&1 | grep -E "<<<<|gimg|SUSPECT|Error|error" | head -40
done > gpurun_out/diag_race_j.log 2>&1
cat gpurun_out/diag_race_j.log
