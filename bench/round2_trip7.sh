#!/bin/bash
# 8-GPU trip (charged 8x): NVLS all-reduce at n = 8, headline bench with the tensor-parallel arm, GPT-2 1.5B pipeline
# (4 stages x 2-way SPMD, 1F1B, CUDA-graph stage bodies), expert-parallel GPT-MoE, data-parallel Wide-ResNet, 4-GPU plan matrix.
out=gpurun_out/r2t7
mkdir -p $out
run8() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run8
step mc_worker         200 bash -c 'run8 29510 tests/mc_worker.py gpurun_out/r2t7/mc.json'
step bench_n8          420 bash -c 'run8 29511 bench.py --gpus 8 --steps 20 --warmup 5'
step bench_pp4m8_1p5B  300 bash -c 'run8 29512 bench.py --gpus 8 --model 1.5B --batch 8 --strategy pp4m8 --steps 10 --warmup 3'
step moe_ep8           200 bash -c 'run8 29513 examples/gpt_moe/train.py --batch 64 --strategy tp --steps 12'
step wrn_dp8           240 bash -c 'run8 29514 examples/wide_resnet/train.py --model-type 1 --batch 32 --steps 20'
step plans_4gpu        400 python -m pytest tests/test_plans_multi_gpu.py -x -q -k four_gpus
cat $out/summary.txt
tail -n 1 $out/bench_n8.log $out/bench_pp4m8_1p5B.log
tail -n 2 $out/moe_ep8.log $out/wrn_dp8.log
