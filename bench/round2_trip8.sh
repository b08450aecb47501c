#!/bin/bash
# 2-GPU trip: bf16 gradient wire over NVLS vs fp32 unicast, own-kernel einsums in GPT-MoE (EP over 2 GPUs), library arm in the
# bench JSON, full GPU test-suite.
out=gpurun_out/r2t8
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
step bench_n2          500 bash -c 'run2 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-tp'
step bench_n2_bf16wire 400 bash -c 'TEPDIST_GRAD_WIRE=bf16 run2 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-tp --no-library-arm'
step moe_ep2           200 bash -c 'run2 29513 examples/gpt_moe/train.py --batch 16 --strategy tp --steps 12'
step moe_ep2_torch     200 bash -c 'TEPDIST_EINSUM=torch run2 29514 examples/gpt_moe/train.py --batch 16 --strategy tp --steps 12'
cat $out/summary.txt
tail -n 1 $out/bench_n2.log $out/bench_n2_bf16wire.log
tail -n 1 $out/moe_ep2.log $out/moe_ep2_torch.log
