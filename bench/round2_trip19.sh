#!/bin/bash
# last GPU seconds of the round: GPU test suite + smoke on the final tree (1 GPU).
out=gpurun_out/r2t19
mkdir -p $out
timeout 60 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $out/summary.txt
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.txt
tail -n 3 $out/pytest_gpu.log; tail -n 1 $out/smoke.log
