#!/bin/bash
# 1-GPU trip: LayerNorm-backward variants (8 warps x 3 CTAs/SM vs 16 warps x 1 CTA/SM), MoE route kernels, racecheck / synccheck.
out=gpurun_out/r2t10
mkdir -p $out
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
step pytest_gpu        900 python -m pytest tests -m gpu -x -q
step moe_routes        200 python tests/kernel_checks.py moe_routes
TEPDIST_PDL=0 step kineto_ln8  150 python bench/kineto_step.py
TEPDIST_PDL=0 TEPDIST_LN_BWD_WARPS=16 step kineto_ln16 150 python bench/kineto_step.py
step bench_n1          300 python bench.py --steps 20 --warmup 5 --no-library-arm
TEPDIST_LN_BWD_WARPS=16 step bench_n1_ln16 300 python bench.py --steps 20 --warmup 5 --no-library-arm
step synccheck         400 bash bench/sanitize.sh synccheck layernorm gelu_colsum_embed gemm_epilogues attn_fwd
step racecheck         500 bash bench/sanitize.sh racecheck layernorm gelu_colsum_embed xent_adam
cat $out/summary.txt
grep "layernorm_bwd" $out/kineto_ln8.log $out/kineto_ln16.log
for f in bench_n1 bench_n1_ln16; do tail -n 1 $out/$f.log | cut -c1-220; done
