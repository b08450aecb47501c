#!/bin/bash
# 1 GPU: ring-attention block schedule + helper kernels, pooling kernels, conv net end to end on the new pooling path.
out=gpurun_out/r2t17
mkdir -p $out
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
step kernels 150 python -m pytest tests/test_kernels_gpu.py -x -q -k "ring_blocks or pool or attn"
step conv1   120 env TEPDIST_TEST_DEVICE=cuda python tests/dist_worker.py conv:auto $out/conv1.json
tail -n 25 $out/kernels.log
tail -n 3 $out/conv1.log
cat $out/conv1.json 2>/dev/null | cut -c1-400
