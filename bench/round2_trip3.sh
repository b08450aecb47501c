#!/bin/bash
# 2-GPU trip: NVLS substrate + multimem kernels, tensor-parallel plan with the fused chains, bench with the TP arm.
out=gpurun_out/r2t3
mkdir -p $out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run2
step mc_worker         300 bash -c 'run2 29510 tests/mc_worker.py gpurun_out/r2t3/mc.json'
TEPDIST_TEST_EXPERIMENTAL=1 step tp_fused_plan 300 python -m pytest tests/test_multi_gpu.py -x -q -k tp_plan
step bench_n2          400 bash -c 'run2 29511 bench.py --gpus 2 --steps 20 --warmup 5'
step gpt2_tp_fused     200 bash -c 'TEPDIST_TP_FUSED=1 run2 29513 examples/gpt2/train.py --model 345M --train-steps 10 --strategy tp'
cat $out/summary.txt
tail -n 1 $out/bench_n2.log
