#!/bin/bash
# final 1-GPU sanity of the tree the driver will run: GPU tests, smoke, bench.
out=gpurun_out/r2t16
mkdir -p $out
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}

step smoke      120 python -c "import __graft_entry__ as g; g.smoke()"
step bench_n1   200 python bench.py --steps 20 --warmup 5
step pytest_gpu 300 python -m pytest tests -m gpu -x -q
cat $out/summary.txt
tail -n 1 $out/bench_n1.log | cut -c1-300
