"""Weight-streaming MoE logit GEMMs (256 x 14148 / 9432 x 4096, bf16 weights from HBM): 128-wide vs 256-wide tiles.
CUDA events around each launch, the L2 flushed (256 MB fill) before every repetition.  EVC_WS128=0/1 selects the rule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efficientvideoclassification_youtube8m_b200 import ops

torch.cuda.set_device(0)
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 4096
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
x = torch.randn(B, S, device="cuda").to(torch.bfloat16)
for name, N in (("gates", 14148), ("experts", 9432)):
    ld = ops.pad8(N, 64)
    W = (torch.randn(S, ld, device="cuda") * 0.02).to(torch.bfloat16)
    out = torch.empty(B, N, device="cuda")
    ts = []
    for rep in range(12):
        ops.fill_f32(flush, 0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(x, W, B, N, S, out, b_mn=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    med = ts[len(ts) // 2]
    print(f"EVC_WS128={os.environ.get('EVC_WS128', '1')} {name:8s} {B}x{N}x{S}: median {med:6.1f} us  min {ts[0]:6.1f} us  "
          f"weights {S * N * 2 / med / 1e6:5.2f} TB/s")

# input-gradient GEMMs of the classifier: dstate[B,S] += dL[B,N] * W[S,N]^T, split-K with f32 reductions into dstate
for name, N in (("gates dgrad", 14148), ("experts dgrad", 9432)):
    ld = ops.pad8(N, 64)
    W = (torch.randn(S, ld, device="cuda") * 0.02).to(torch.bfloat16)
    dl = torch.zeros(B, ld, device="cuda", dtype=torch.bfloat16)
    dl[:, :N] = (torch.randn(B, N, device="cuda") * 1e-2).to(torch.bfloat16)
    out = torch.zeros(B, S, device="cuda")
    for sk in (2, 4, 6, 8, 9):
        ts = []
        for rep in range(12):
            ops.fill_f32(flush, 0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(dl, W, B, S, N, out, split_k=sk, accumulate=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        print(f"{name:14s} {B}x{S}x{N} split_k={sk}: median {ts[len(ts) // 2]:6.1f} us  min {ts[0]:6.1f} us")
