import sys, ctypes; sys.path.insert(0, '.')
import torch
from efficientvideoclassification_youtube8m_b200 import ops, _lib
dbg = ctypes.CDLL(_lib.LIB_PATH).evc_debug_set
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(76800, 1024, 4096), (5120, 4096, 2176), (5120, 1024, 4096), (1152, 4096, 76800)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda")
    res = []
    for flag in [0, 250<<16, 500<<16, 1000<<16, 2000<<16, 32]:
        dbg(flag)
        res.append(f"dbg{flag>>16 if flag>>16 else flag} {timeit(lambda: ops.gemm(A, B, M, N, K, out)):7.1f}us")
    dbg(0)
    print((M, N, K), " | ".join(res), flush=True)
