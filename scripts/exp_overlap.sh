#!/bin/bash
# Stream-overlap ablation: one short bench per EVC_OVERLAP mode (and stream priorities), same box.
mkdir -p gpurun_out
out=gpurun_out/exp_overlap.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" python bench.py --skip-cpu --skip-infer --steps 20 --warmup 3 2>>gpurun_out/exp_overlap.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('ms/step %.3f  value %.0f  e2e %.0f  u8 %.0f  clocks %s  loss %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_uint8_input']['value'], d['clocks']['sm_mhz'], {k: round(v, 4) for k, v in d['losses'].items()}))
" >> $out
}
run EVC_OVERLAP=0
run EVC_OVERLAP=7
run EVC_OVERLAP=1
run EVC_OVERLAP=2
run EVC_OVERLAP=4
run EVC_OVERLAP=5
run EVC_OVERLAP=7 EVC_PRIO_STUDENT=-1
run EVC_OVERLAP=7 EVC_PRIO_SIDE=-1
run EVC_OVERLAP=0
cat $out
