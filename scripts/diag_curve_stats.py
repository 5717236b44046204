import sys, json, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, './tests')
import test_gpu_parity as T
floor = T._floor(1e-3)["bf16"]["summary"]
r4 = T._curve(1e-4, 200)
print("lr1e-4 max", {k: round(float(v.max()), 5) for k, v in r4.items()})
for rep in range(4):
    rel = T._curve(1e-3, 200)
    print("lr1e-3 rep", rep, {k: (round(float(np.percentile(v, 95)), 4), round(float(v.max()), 4), round(float(np.median(v)), 4),
                                  "floor p95 %.4f max %.4f" % (floor[k]["p95"], floor[k]["max"])) for k, v in rel.items()})
