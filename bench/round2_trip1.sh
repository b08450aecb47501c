#!/bin/bash
# First GPU trip of the next round: ONE gpurun call (1 GPU, ~8-10 min of box time) that re-validates the default path and
# works through the queue of code that was written after the GPU budget of round 1 was spent (NEXT_STEPS.md).
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash bench/round2_trip1.sh'
#
# Every step has its own timeout and log under gpurun_out/r2t1/ ; a failing or hanging step does not stop the others.
# Nothing here runs under a profiler, so the bench numbers are usable.  Copy what should be judged into profiles/.
out=gpurun_out/r2t1
mkdir -p $out
step() {   # step <name> <timeout-s> <command...>
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/gpu.txt 2>&1

# 1. what the driver runs at round end (default path edits since the last hardware run: NEXT_STEPS "edited AFTER")
step pytest_gpu        600 python -m pytest tests -m gpu -x -q
step smoke             120 python -c "import __graft_entry__ as g; g.smoke()"
step bench_n1          180 python bench.py --steps 20 --warmup 5

# 2. in-situ kernel breakdown of the captured step (never run so far)
step kineto            150 python bench/kineto_step.py

# 3. prepared-but-never-executed kernels (default off)
step attn_poly         120 python tests/kernel_checks.py attn_poly
step attn_fwd2         120 python tests/kernel_checks.py attn_fwd2
TEPDIST_ATTN_EXP_POLY=1 step bench_n1_exp_poly 150 python bench.py --steps 20 --warmup 5 --no-exposed

# 4. optimizers added on CPU only: CUDA-graph capture must not sync, losses must fall
for opt in adafactor lamb sm3; do
  step gpt2_117M_$opt  150 python examples/gpt2/train.py --model 117M --train-steps 6 --optimizer $opt
done
step moe_tiny_adafactor 120 python examples/gpt_moe/train.py --tiny --batch 4 --steps 4 --optimizer adafactor
step torch_frontend     120 python examples/torch_frontend/train.py
step run_graph_tool     120 python -m tepdist_b200.tools.run_graph --snippet attention --profile

cat $out/summary.txt
tail -n 2 $out/bench_n1.log $out/bench_n1_exp_poly.log
