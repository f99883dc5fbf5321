#!/bin/bash
# ncu --set full over the forward convolutions of one step (32 frames, native engine, 3xTF32) + the stem dgrad;
# summaries are extracted ON THE BOX (the .ncu-rep itself can exceed the 64 MiB return limit)
mkdir -p gpurun_out /tmp/prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_persist|stem_" -s 24 -c 26 -o /tmp/prof/conv python bench.py --steps 1 --warmup 1 --clips 1 --engine native --no-cpu-baseline --no-e2e > gpurun_out/ncu_k.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stem_dgrad" -s 0 -c 1 -o /tmp/prof/stemd python bench.py --steps 1 --warmup 1 --clips 1 --engine native --no-cpu-baseline --no-e2e >> gpurun_out/ncu_k.log 2>&1
ls -la /tmp/prof
ncu -i /tmp/prof/conv.ncu-rep --page raw --csv > gpurun_out/conv_raw.csv 2>/dev/null
ncu -i /tmp/prof/stemd.ncu-rep --page raw --csv > gpurun_out/stemd_raw.csv 2>/dev/null
ncu -i /tmp/prof/stemd.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/stemd_source.csv.gz
# source pages: launch 1 = stem fwd, 2.. = tc kernels in forward order (l1.0.ds, l1.0.conv1, conv2, conv3, ...)
for i in 0 1 2 3 4 12; do
  ncu -i /tmp/prof/conv.ncu-rep --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | gzip > gpurun_out/conv_source_$i.csv.gz
done
ls -la gpurun_out | tail -12
