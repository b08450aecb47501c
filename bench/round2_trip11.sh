#!/bin/bash
# 2-GPU trip: plan matrix incl. native synchronised BatchNorm, bench (stream-K pool fix, library arm), bf16 gradient wire over
# NVLS, GPT-MoE with route-table dispatch / combine under a CUDA graph vs dense einsums, Wide-ResNet DP with native sync-BN.
out=gpurun_out/r2t11
mkdir -p $out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run2
step plans_multi_gpu   600 python -m pytest tests/test_plans_multi_gpu.py -x -q -k two_gpus
step multi_gpu_tests   600 python -m pytest tests/test_multi_gpu.py -x -q
step bench_n2          500 bash -c 'run2 29511 bench.py --gpus 2 --steps 20 --warmup 5'
step bench_n2_bf16wire 400 bash -c 'TEPDIST_GRAD_WIRE=bf16 run2 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-tp --no-library-arm'
step moe_sparse_graph  200 bash -c 'run2 29513 examples/gpt_moe/train.py --batch 16 --strategy tp --steps 12'
step moe_dense_graph   200 bash -c 'TEPDIST_MOE_SPARSE=0 run2 29514 examples/gpt_moe/train.py --batch 16 --strategy tp --steps 12'
step moe_sparse_eager  200 bash -c 'run2 29515 examples/gpt_moe/train.py --batch 16 --strategy tp --steps 12 --no-graph'
step wrn_dp2           240 bash -c 'run2 29516 examples/wide_resnet/train.py --model-type 1 --batch 8 --steps 20'
step wrn_dp2_torchbn   240 bash -c 'TEPDIST_BN_SYNC=torch run2 29517 examples/wide_resnet/train.py --model-type 1 --batch 8 --steps 20'
cat $out/summary.txt
tail -n 1 $out/bench_n2.log | cut -c1-400
tail -n 1 $out/bench_n2_bf16wire.log | cut -c1-300
tail -n 1 $out/moe_sparse_graph.log $out/moe_dense_graph.log $out/moe_sparse_eager.log $out/wrn_dp2.log $out/wrn_dp2_torchbn.log
