#!/bin/bash
# Second GPU trip of the next round: 2 GPUs (charged 2x), ~6 min of box time.
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 700 -- 'bash bench/round2_trip2.sh'
#
out=gpurun_out/r2t2
mkdir -p $out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run2
step mc_worker         300 bash -c 'run2 29510 tests/mc_worker.py gpurun_out/r2t2/mc.json'
step pytest_multi_gpu  500 python -m pytest tests/test_multi_gpu.py -x -q
TEPDIST_TEST_EXPERIMENTAL=1 step tp_fused_plan 200 python -m pytest tests/test_multi_gpu.py -x -q -k tp_plan
step bench_n2          200 bash -c 'run2 29511 bench.py --gpus 2 --steps 20 --warmup 5'
# pipeline with the persistent receive ring + scheduler-driven release (only gloo-tested so far), then the tp plan after the
# planner fix, then a sharded-update optimizer with cross-rank reductions over NCCL
step gpt2_pp2m4        200 bash -c 'run2 29512 examples/gpt2/train.py --model 345M --train-steps 10 --strategy pp2m4'
step gpt2_tp           200 bash -c 'run2 29513 examples/gpt2/train.py --model 345M --train-steps 10 --strategy tp --comm nccl'
step gpt2_adafactor_dp 200 bash -c 'run2 29514 examples/gpt2/train.py --model 117M --train-steps 6 --optimizer adafactor --comm nccl'
step moe_ep            200 bash -c 'run2 29515 examples/gpt_moe/train.py --steps 10'
cat $out/summary.txt
