#!/bin/bash
# 4-GPU A/B: TEPDIST_OVERLAP_CTAS 296 (default so far) vs 74, alternating, data-parallel step only.
out=gpurun_out/r2t15
mkdir -p $out
run4() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
export -f run4
i=0
for c in 296 74 296 74 37; do
  i=$((i+1))
  TEPDIST_OVERLAP_CTAS=$c timeout 200 bash -c "run4 $((29520+i)) bench.py --gpus 4 --steps 30 --warmup 5 --no-tp --no-library-arm --no-exposed" > $out/run${i}_ctas$c.log 2>&1
  echo "run$i ctas=$c rc=$? $(tail -n 1 $out/run${i}_ctas$c.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>/dev/null)" | tee -a $out/summary.txt
done
cat $out/summary.txt
