#!/bin/bash
# 4-GPU trip: the final tree on the driver's command (bench with the TP and library arms), overlap-CTA sweep for the exposed
# communication of the data-parallel step, multi-GPU test-suites.
out=gpurun_out/r2t14
mkdir -p $out
run4() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
step() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$out/$name.log" 2>&1
  echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a $out/summary.txt
}
export -f run4
step bench_n4          500 bash -c 'run4 29511 bench.py --gpus 4 --steps 20 --warmup 5'
step bench_n4_ctas148  300 bash -c 'TEPDIST_OVERLAP_CTAS=148 run4 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-tp --no-library-arm'
step bench_n4_ctas74   300 bash -c 'TEPDIST_OVERLAP_CTAS=74 run4 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-tp --no-library-arm'
step bench_n4_first2m  300 bash -c 'TEPDIST_FIRST_BUCKET=2097152 run4 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-tp --no-library-arm'
step plans_multi_gpu   600 python -m pytest tests/test_plans_multi_gpu.py -x -q
step multi_gpu_tests   600 python -m pytest tests/test_multi_gpu.py -x -q
cat $out/summary.txt
for f in bench_n4 bench_n4_ctas148 bench_n4_ctas74 bench_n4_first2m; do tail -n 1 $out/$f.log | cut -c1-420; done
