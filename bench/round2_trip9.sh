#!/bin/bash
# 1-GPU trip: stream-K workspace pool fix, LayerNorm-backward cluster reduction, library arm in the bench JSON, einsum check.
out=gpurun_out/r2t9
mkdir -p $out
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
step pytest_gpu        900 python -m pytest tests -m gpu -x -q
step bench_n1          300 python bench.py --steps 20 --warmup 5
step einsum            200 python tests/kernel_checks.py einsum
TEPDIST_PDL=0 step kineto_nopdl 150 python bench/kineto_step.py
step smoke             120 python -c "import __graft_entry__ as g; g.smoke()"
cat $out/summary.txt
tail -n 1 $out/bench_n1.log | cut -c1-1500
