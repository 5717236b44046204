#!/bin/bash
# 8-GPU A/B of the early classifier optimizer/all-gather (EVC_EARLY_OPT) and the final bench line
N=${1:-8}
OUT=gpurun_out/r02_dp_early_ab_${N}gpu.txt
: > $OUT
for eo in 1 0 1 0; do
  EVC_EARLY_OPT=$eo python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --steps 20 --warmup 3 --skip-f32-e2e --skip-tfrecord > gpurun_out/_ab.json 2> gpurun_out/_ab.err
  python - <<PY >> $OUT
import json
try:
    d = json.loads(open("gpurun_out/_ab.json").read().strip().splitlines()[-1])
    print("EVC_EARLY_OPT=$eo n_gpus", d["n_gpus"], "ms_per_step %.3f" % d["ms_per_step"], "videos/s %.0f" % d["value"],
          "e2e ms %.3f" % d["e2e"]["ms_per_step"], "e2e videos/s %.0f" % d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("EVC_EARLY_OPT=$eo failed:", e, open("gpurun_out/_ab.err").read()[-800:])
PY
done
cat $OUT
