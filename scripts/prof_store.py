import sys, ctypes; sys.path.insert(0, '.')
import torch
from efficientvideoclassification_youtube8m_b200 import ops, _lib
flag = int(sys.argv[1])
M, N, K = 5120, 4096, 2176
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
out = torch.zeros(M, N, device="cuda")
_lib.lib.evc_debug_set(flag)
for _ in range(5): ops.gemm(A, B, M, N, K, out)
torch.cuda.synchronize()
