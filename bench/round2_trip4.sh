#!/bin/bash
# 4-GPU trip: plan matrix on NCCL (tp / pipeline / pipeline x SPMD / 2-D mesh / expert parallel), NVLS kernels at n = 4,
# bench with the TP arm, pipeline bench (345M, 2 stages x 2-way SPMD).
out=gpurun_out/r2t4
mkdir -p $out
run4() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run4
step plans_multi_gpu   600 python -m pytest tests/test_plans_multi_gpu.py -x -q
step mc_worker         300 bash -c 'run4 29510 tests/mc_worker.py gpurun_out/r2t4/mc.json'
step bench_n4          500 bash -c 'run4 29511 bench.py --gpus 4 --steps 20 --warmup 5'
step bench_n4_unicast  300 bash -c 'TEPDIST_DP_MC=0 run4 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-tp'
step bench_pp2m8       300 bash -c 'run4 29513 bench.py --gpus 4 --steps 10 --warmup 3 --strategy pp2m8'
step multi_gpu_tests   600 python -m pytest tests/test_multi_gpu.py -x -q
cat $out/summary.txt
tail -n 1 $out/bench_n4.log $out/bench_n4_unicast.log $out/bench_pp2m8.log
