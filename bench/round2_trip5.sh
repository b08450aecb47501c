#!/bin/bash
# 1-GPU trip: TMA-store epilogue of the 2-CTA GEMM, A/B on one box (numerics, per-shape TFLOP/s vs cuBLAS, step time).
out=gpurun_out/r2t5
mkdir -p $out
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
step gemm2_tma         200 python tests/kernel_checks.py gemm2
cp gpurun_out/kernel_checks.json $out/kernel_checks_tma.json 2>/dev/null
TEPDIST_GEMM2_TMA_STORE=0 step gemm2_reg 200 python tests/kernel_checks.py gemm2
cp gpurun_out/kernel_checks.json $out/kernel_checks_reg.json 2>/dev/null
step pytest_gpu        600 python -m pytest tests -m gpu -x -q
step bench_tma         200 python bench.py --steps 20 --warmup 5
TEPDIST_GEMM2_TMA_STORE=0 step bench_reg 200 python bench.py --steps 20 --warmup 5
step bench_tma2        200 python bench.py --steps 20 --warmup 5
TEPDIST_PDL=0 step kineto_nopdl 150 python bench/kineto_step.py
cat $out/summary.txt
for f in bench_tma bench_reg bench_tma2; do tail -n 1 $out/$f.log | cut -c1-200; done
