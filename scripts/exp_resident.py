"""Micro-benchmark of the small-row recurrence paths: resident-weights persistent kernel vs per-step split-K GEMM +
cell kernel (CUDA events, 20 repetitions after 3 warm-ups)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efficientvideoclassification_youtube8m_b200 import _lib, ops

def bench(fn, n=20, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

reps = int(os.environ.get("REPS", "20"))
for rows, Kx, H, T in [(256, 4096, 1024, 20), (256, 1024, 1024, 20), (256, 4096, 1024, 5), (256, 1024, 1024, 5),
                       (512, 4096, 1024, 5), (1280, 1152, 1024, 6), (1024, 4096, 1024, 5)]:
    dev = "cuda"
    x = (torch.randn(T, rows, Kx, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(Kx + H, 4 * H, device=dev) * (2.0 / (Kx + H) ** 0.5)).to(torch.bfloat16)
    b = torch.randn(4 * H, device=dev) * 0.1
    seq_len = torch.full((rows,), T, device=dev, dtype=torch.int32)
    h_all = torch.zeros(T + 1, rows, H, dtype=torch.bfloat16, device=dev)
    c_all = torch.zeros(T + 1, rows, H, device=dev)
    gates = torch.zeros(T, rows, 4 * H, dtype=torch.bfloat16, device=dev)
    need = ops.lstm_rec_workspace_bytes(rows, H, T)
    raw = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
    ws_rec = raw[(-raw.data_ptr()) % 1024:][:need]
    ws = torch.empty(ops.lstm_workspace_bytes(rows, H, Kx), dtype=torch.uint8, device=dev)
    zx = torch.empty(T * rows, 4 * H, device=dev)
    t_res = bench(lambda: ops.lstm_seq_fwd_resident(x, rows * Kx, Kx, W, b, rows, H, T, seq_len, h_all, c_all, gates, ws_rec), reps)
    t_step = bench(lambda: ops.lstm_seq_fwd(x, rows * Kx, Kx, W, b, rows, H, T, seq_len, h_all, c_all, gates, ws), reps)
    _lib.lib.evc_debug_set(8192)     # slab path: split-K GEMM + cell kernel
    t_slab = bench(lambda: ops.lstm_seq_fwd(x, rows * Kx, Kx, W, b, rows, H, T, seq_len, h_all, c_all, gates, ws), reps)
    _lib.lib.evc_debug_set(0)
    t_zx = bench(lambda: ops.gemm(x.view(T * rows, Kx), W, T * rows, 4 * H, Kx, zx, b_mn=True, ldb=4 * H, bias=b), reps)
    print(f"rows {rows} Kx {Kx} H {H} T {T}: resident {t_res:8.1f} us (of which Zx GEMM {t_zx:7.1f}) -> {(t_res - t_zx) / T:6.1f} us/step rec;"
          f"  per-step default (cluster split-K when eligible) {t_step:8.1f} us = {t_step / T:6.1f} us/step;"
          f"  slab path {t_slab:8.1f} us = {t_slab / T:6.1f} us/step")
