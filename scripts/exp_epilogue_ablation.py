"""What does each part of the fused LSTM forward epilogue cost the main loop?  

Builds a second library from the same sources with -DEVC_ABLATE (the product library contains none of the
hooks), re-executes itself with EVC_LIB_PATH pointing at it, and times the 15 forward steps of the teacher's
RNN_L1 cell 0 (5120 rows, K = 2176, the bench's dominant kernel) by CUDA events for each combination of
hooks.  The outputs of the ablated runs are garbage by construction; only the times mean anything.

    python scripts/exp_epilogue_ablation.py            # on a B200
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "efficientvideoclassification_youtube8m_b200", "csrc")
ABL = os.path.join(ROOT, "gpurun_out", "libevc_ablate.so")

CASES = [
    (0, "full epilogue"),
    (32, "no tcgen05.ld (main loop + barriers only)"),
    (1, "tcgen05.ld only"),
    (4 | 8 | 16, "ld + transposition + math, no global I/O"),
    (2 | 4 | 8 | 16, "ld + math, no shared-memory stores, no global I/O"),
    (8 | 16, "everything but the stores"),
    (4, "everything but the global loads"),
    (8, "everything but the gate stores"),
    (16, "everything but the c/h stores"),
]


def build():
    os.makedirs(os.path.dirname(ABL), exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DEVC_ABLATE",
           "-shared", "-Xcompiler", "-fPIC", "-o", ABL] + [os.path.join(CSRC, f) for f in
                                                            ("evc_host.cu", "evc_kernels.cu", "evc_gemm.cu")]
    subprocess.run(cmd, check=True, cwd=CSRC)


def main():
    if os.environ.get("EVC_LIB_PATH") != ABL:
        if "--no-build" not in sys.argv:
            build()
        os.execve(sys.executable, [sys.executable] + sys.argv, dict(os.environ, EVC_LIB_PATH=ABL, EVC_OVERLAP="0"))
    import torch
    from efficientvideoclassification_youtube8m_b200 import _lib, ops
    from efficientvideoclassification_youtube8m_b200 import synthetic as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    B, cfg = 256, ModelConfig()
    x, nf, lab = O.synthetic_batch(B, seed=1234, full_length=True)
    tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda", base_learning_rate=1e-5)
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
    tr.step(xd, nfd, labd)
    torch.cuda.synchronize()
    t, p = tr.t_eng, tr.teacher
    H, D, R1, ell = cfg.lstm_cells, cfg.feature_size, t.R1, t.ell
    lay = t.l1[0]

    def l1_fwd():
        ops.lstm_seq_fwd(t.x, R1 * D, D, p.shadow[p.kernel(0, 0)], p.w[p.bias(0, 0)], R1, H, ell, t.len_l1,
                         lay.h_all, lay.c_all, lay.gates)

    def timed(reps=5):
        for _ in range(2):
            l1_fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            l1_fwd()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * ell)

    print(f"{'us/step':>8}  hooks  what")
    for rnd in range(2):                      # twice: clocks settle under the power cap
        for flags, what in CASES:
            _lib.lib.evc_debug_set(flags)
            print(f"{timed():8.1f}  {flags:5d}  {what}", flush=True)
    _lib.lib.evc_debug_set(0)


if __name__ == "__main__":
    main()
