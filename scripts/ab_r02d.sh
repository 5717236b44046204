#!/bin/bash
# 1-GPU A/B in one call (same box): stream-overlap modes with the round's final kernels
OUT=gpurun_out/r02_ab_overlap_final.txt
: > $OUT
run() {
  env "$@" python bench.py --steps 20 --warmup 3 --skip-cpu --skip-tfrecord --skip-configs --skip-f32-e2e --skip-infer \
    > gpurun_out/_ab.json 2> gpurun_out/_ab.err
  python - "$*" <<'PY' >> gpurun_out/r02_ab_overlap_final.txt
import json, sys
try:
    d = json.loads(open("gpurun_out/_ab.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "| ms_per_step %.3f" % d["ms_per_step"], "videos/s %.0f" % d["value"],
          "e2e ms %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "SM MHz", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "failed:", e, open("gpurun_out/_ab.err").read()[-1500:])
PY
}
for rep in 1 2; do
  run EVC_OVERLAP=7
  run EVC_OVERLAP=15
  run EVC_OVERLAP=7 EVC_FUSED_NORMS=0
  run EVC_OVERLAP=3
done
cat $OUT
