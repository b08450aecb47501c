#!/bin/bash
# 1-GPU trip: ncu --set full captures of the final kernels (2-CTA GEMM with the TMA-store epilogue in its three hot
# instantiations, LayerNorm fwd / bwd, MoE route kernels + expert FC einsum), summaries written next to the reports.
out=gpurun_out/r2t13
mkdir -p $out
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
NCU="ncu --set full --clock-control none --import-source on"
step ncu_gemm2 300 $NCU -k regex:gemm2_bf16 -s 3 -c 6 -f -o $out/prof_gemm2_final python bench/prof_gemm.py gemm2_final
step ncu_ln    200 $NCU -k regex:layernorm -s 2 -c 4 -f -o $out/prof_ln python bench/prof_gemm.py ln
step ncu_moe   300 $NCU -k regex:moe_\|gemm_bf16 -s 4 -c 8 -f -o $out/prof_moe python bench/prof_gemm.py moe
step ncu_attn  300 $NCU -k regex:attn_ -s 4 -c 4 -f -o $out/prof_attn python bench/prof_gemm.py attn
step pytest_gpu 900 python -m pytest tests -m gpu -x -q
step bench_n1  300 python bench.py --steps 20 --warmup 5
cat $out/summary.txt
tail -n 1 $out/bench_n1.log | cut -c1-300
