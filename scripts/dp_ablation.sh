#!/bin/bash
# Attribution of the multi-GPU step time (no nsys in this image): the same N-GPU bench with the gradient collectives
# (EVC_DP_ABLATE=1), the operand all-gathers (=2) or both (=3) switched off.  The ablated runs compute garbage across
# ranks but keep the kernel schedule, so the differences are the exposed cost of each collective family.
N=${1:-8}
OUT=gpurun_out/r02_dp_ablation_${N}gpu.txt
: > $OUT
for mode in 0 1 2 3; do
  EVC_DP_ABLATE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 --skip-f32-e2e --skip-tfrecord > gpurun_out/_abl.json 2> gpurun_out/_abl.err
  python - <<PY >> $OUT
import json
try:
    d = json.loads(open("gpurun_out/_abl.json").read().strip().splitlines()[-1])
    print("EVC_DP_ABLATE=$mode n_gpus", d["n_gpus"], "ms_per_step %.3f" % d["ms_per_step"], "videos/s %.0f" % d["value"],
          "e2e ms %.3f" % d["e2e"]["ms_per_step"], "e2e videos/s %.0f" % d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("EVC_DP_ABLATE=$mode failed:", e, open("gpurun_out/_abl.err").read()[-800:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
cat $OUT
