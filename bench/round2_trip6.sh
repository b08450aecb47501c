#!/bin/bash
# 2-GPU trip: plan matrix over NCCL, pipeline with CUDA-graph stage bodies (A/B against eager), sanitizer pass on the kernels.
out=gpurun_out/r2t6
mkdir -p $out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run2
step plans_multi_gpu   600 python -m pytest tests/test_plans_multi_gpu.py -x -q
step bench_pp2m8_graph 300 bash -c 'run2 29513 bench.py --gpus 2 --steps 10 --warmup 3 --strategy pp2m8'
step bench_pp2m8_eager 300 bash -c 'TEPDIST_PP_GRAPH=0 run2 29514 bench.py --gpus 2 --steps 10 --warmup 3 --strategy pp2m8'
step pp_example        200 bash -c 'run2 29515 examples/gpt2/train.py --model 345M --train-steps 8 --strategy pp2m4'
step sanitize_memcheck 500 bash bench/sanitize.sh memcheck
cat $out/summary.txt
tail -n 1 $out/bench_pp2m8_graph.log $out/bench_pp2m8_eager.log
