#!/bin/bash
# 2 GPUs: context-parallel GPT-2 (sequence 256 = 2 blocks of 128, ring attention over NCCL p2p) vs the data-parallel plan of the same model.
out=gpurun_out/r2t18
mkdir -p $out
export TEPDIST_TEST_DEVICE=cuda TEPDIST_TEST_NCTX=256 OMP_NUM_THREADS=2
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29731 \
  tests/dist_worker.py "gpt2:cp+gpt2:auto" $out/cp2.json > $out/cp2.log 2>&1
echo "rc=$?" | tee $out/summary.txt
grep -v "^W0\|^\*\*\*" $out/cp2.log | tail -n 12
python - <<'PY' | tee -a gpurun_out/r2t18/summary.txt
import json
d = json.load(open("gpurun_out/r2t18/cp2.json"))
for k, v in d.items():
    print(k, v["parallelism"], [round(x, 4) for x in v["losses"]], v.get("collectives"))
a, b = d["gpt2:cp"]["losses"], d["gpt2:auto"]["losses"]
print("max rel diff", max(abs(x - y) / abs(y) for x, y in zip(a, b)))
PY
