"""Yardstick: this repo's tcgen05 GEMM vs torch.matmul (cuBLAS) on the path's shapes (timing only)."""
import sys; sys.path.insert(0, '.')
import torch
from efficientvideoclassification_youtube8m_b200 import ops
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(5120, 1024, 4096), (5120, 4096, 2176), (76800, 1024, 4096), (256, 4096, 5120), (256, 1024, 4096), (1280, 1024, 4096), (1152, 4096, 76800)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    Bt = B.t().contiguous()
    out = torch.zeros(M, N, device="cuda")
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t_cublas = timeit(lambda: torch.matmul(A, Bt))
    res = [f"cublas(bf16 out) {t_cublas:7.1f}us {2*M*N*K/t_cublas/1e6:7.1f}TF"]
    for sk in [1, 2, 3, 4, 8]:
        if K // 64 // sk < 4: continue
        t = timeit(lambda: ops.gemm(A, B, M, N, K, out, split_k=sk))
        res.append(f"sk{sk} {t:7.1f}us")
    t = timeit(lambda: ops.gemm(A, B, M, N, K, outb))
    res.append(f"bf16out {t:7.1f}us")
    t = timeit(lambda: ops.gemm(A, Bt, M, N, K, out, b_mn=True))
    res.append(f"Bmn {t:7.1f}us")
    print((M, N, K), " | ".join(res), flush=True)
