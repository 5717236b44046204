#!/bin/bash
# Schedule ablation: one short bench per setting, same box.
mkdir -p gpurun_out
out=gpurun_out/exp_overlap3.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" python bench.py --skip-cpu --skip-infer --steps 20 --warmup 3 2>>gpurun_out/exp_overlap.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('ms/step %.3f  value %.0f  e2e %.0f  u8 %.0f  clocks %s  seq_fwd_us %.1f loss %.4f %.4f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_uint8_input']['value'], d['clocks']['sm_mhz'], d['roofline']['avg_launch_ms']*1e3, d['losses']['teacher_loss'], d['losses']['student_loss']))
" >> $out
}
run EVC_OVERLAP=15
run EVC_OVERLAP=7
run EVC_OVERLAP=15 EVC_DGRAD_BN=128
run EVC_OVERLAP=0 EVC_DGRAD_BN=128
run EVC_OVERLAP=0
run EVC_OVERLAP=15
cat $out
