#!/bin/bash
mkdir -p gpurun_out
EVC_OVERLAP=23 timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for m in 7 23 7 23; do
  echo "== EVC_OVERLAP=$m"
  EVC_OVERLAP=$m python bench.py --skip-cpu --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f value %.0f e2e %.0f clocks %s infer %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], {k: round(v, 1) for k, v in d['student_infer'].items()}))"
done
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
python bench.py --workload finetune_cfg4 --skip-cpu --steps 20 > gpurun_out/bench_final_cfg4.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01_s2.csv python scripts/profile_step.py 3 > gpurun_out/launches_r01_s2.log 2>&1
wc -l gpurun_out/launches_r01_s2.csv
