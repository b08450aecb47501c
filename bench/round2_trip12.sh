#!/bin/bash
# 2-GPU trip: GPT-MoE route-table path after the sharding guard (graph + eager), Wide-ResNet DP with a CUDA graph, new kernel checks.
out=gpurun_out/r2t12
mkdir -p $out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run2
step pytest_gpu        900 python -m pytest tests -m gpu -x -q
step moe_sparse_graph  200 bash -c 'run2 29513 examples/gpt_moe/train.py --batch 16 --strategy tp --steps 12'
step moe_dense_graph   200 bash -c 'TEPDIST_MOE_SPARSE=0 run2 29514 examples/gpt_moe/train.py --batch 16 --strategy tp --steps 12'
step moe_sparse_1gpu   200 python examples/gpt_moe/train.py --batch 8 --steps 12
step moe_dense_1gpu    200 env TEPDIST_MOE_SPARSE=0 python examples/gpt_moe/train.py --batch 8 --steps 12
step wrn_dp2_graph     240 bash -c 'run2 29516 examples/wide_resnet/train.py --model-type 1 --batch 8 --steps 20 --graph'
step wrn_dp2_eager     240 bash -c 'run2 29517 examples/wide_resnet/train.py --model-type 1 --batch 8 --steps 20'
step wrn_1gpu_graph    240 python examples/wide_resnet/train.py --model-type 1 --batch 4 --steps 20 --graph
cat $out/summary.txt
tail -n 1 $out/moe_sparse_graph.log $out/moe_dense_graph.log $out/moe_sparse_1gpu.log $out/moe_dense_1gpu.log $out/wrn_dp2_graph.log $out/wrn_dp2_eager.log $out/wrn_1gpu_graph.log
