#!/bin/bash
# last check of HEAD: full GPU suite + smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/tests_av.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_av.log
tail -4 gpurun_out/tests_av.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
