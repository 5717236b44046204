#!/bin/bash
# N-GPU A/B (same box): classifier weight gradients from the gathered batch (EVC_DP_GATHER=1, default) vs reduce-scatter
N=${1:-2}
OUT=gpurun_out/r02_dp_gather_ab_${N}gpu.txt
: > $OUT
for g in ${GATHER_ORDER:-1 0 1 0}; do
  EVC_DP_GATHER=$g python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --steps 20 --warmup 3 --skip-f32-e2e --skip-tfrecord --skip-configs --skip-infer > gpurun_out/_ab.json 2> gpurun_out/_ab.err
  python - <<PY >> $OUT
import json
try:
    d = json.loads(open("gpurun_out/_ab.json").read().strip().splitlines()[-1])
    print("EVC_DP_GATHER=$g n_gpus", d["n_gpus"], "ms_per_step %.3f" % d["ms_per_step"], "videos/s %.0f" % d["value"],
          "e2e ms %.3f" % d["e2e"]["ms_per_step"], "e2e videos/s %.0f" % d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"],
          "losses", {k: round(v, 3) for k, v in d["losses"].items() if k in ("teacher_loss", "student_loss")})
except Exception as e:
    print("EVC_DP_GATHER=$g failed:", e, open("gpurun_out/_ab.err").read()[-1500:])
PY
done
cat $OUT
