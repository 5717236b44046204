#!/usr/bin/env python
"""Benchmark of the H-LSTM teacher-student hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference graph

A "step" is one joint teacher+student training iteration (run_train.sh defaults: batch 256 per
GPU, 300 frames x 1152-d, every_n=10, 2x1024 LSTM cells, 2 mixtures, 4716 classes) on synthetic
inputs.  One JSON line is printed by rank 0 (see the driver contract in the task statement).

  value    videos/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e      videos/s through the public step API with the quantised uint8 batch (what the tfrecords hold) copied
           from pinned host memory every step (double buffered on a copy stream), Dequantize fused into the pack
           kernel, and the losses + top-20 read back
  e2e_f32_input / e2e_tfrecord_input   the same with the dequantised float32 batch (4x the bytes), and with the
           batch decoded from a TFRecord shard on disk by the native reader inside the timed loop (every rank its own)
  configs  BASELINE configs #4 (fine-tune, 2048 cells x 4 mixtures), #5 (every_n sweep, random vs uniform) and the
           split-bf16 precise mode, a few steps each (one GPU)
  roofline dominant kernel = fused LSTM forward step GEMM of the teacher's lower level, timed alone
  cpu_baseline  the float32 PyTorch-CPU restatement (oracle/, "port": TensorFlow 1.x is not
           installable here) on a bounded sample, on the host's cores
"""
from __future__ import annotations

import argparse
import json

import numpy as np
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GF_PER_VIDEO_TRAIN = 39.820  # T+S train step, algorithmic GFLOP per video (cfg #1)
GF_PER_VIDEO_FINETUNE_CFG4 = 16.368  # student fine-tune step at H=2048, M=4 (cfg #4)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # median over the samples taken under load (upper half: idle samples at the edges are lower)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(batch: int, steps: int, warmup: int, finetune: bool = False):
    """float32 CPU restatement of the same graph (oracle, kind 'port'), all host threads."""
    import torch
    from oracle import hlstm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    x, nf, lab = O.synthetic_batch(batch, seed=1234, full_length=True)
    T = O.init_params("model", 0, dtype=torch.float32)
    S = O.init_params("model_student", 1, dtype=torch.float32)
    ot, os_ = O.TFAdam(T), O.TFAdam(S)
    xt, lt = torch.from_numpy(x), torch.from_numpy(lab)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.teacher_student_train_step(xt, nf, lt, T, S, ot, os_)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch / sec, sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # 32 videos per step: large enough to amortise the batch-independent optimizer pass over 286 M parameters
    # (a 4-video sample under-reports the CPU path 6x), small enough for K steps within minutes
    sample_batch = 32
    vps, sec, cores = cpu_baseline(sample_batch, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "H-LSTM teacher-student train videos/s", "value": vps, "unit": "videos/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "teacher-student joint train step (run_train.sh defaults), CPU restatement of the "
                               "reference TF graph (TensorFlow 1.x unavailable offline)",
                   "batch_per_step": sample_batch, "frames": 300, "features": 1152, "every_n": 10,
                   "lstm_cells": 1024, "lstm_layers": 2, "moe_num_mixtures": 2, "classes": 4716},
        "cpu_baseline": {"value": vps, "unit": "videos/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of a {sample_batch}-video batch (same graph, f32)"},
        "e2e": {"value": vps, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _tfrecord_e2e(tr, B, dev, steps, ops, rank=0):
    """videos/s of shard files -> reader -> GPU step -> host results (one shard of B synthetic videos per rank,
    read `steps + warm-up` times; the page cache holds it, as it would hold a training set's hot shards)."""
    import tempfile
    import torch
    from efficientvideoclassification_youtube8m_b200 import readers as R
    rng = np.random.default_rng(77 + rank)
    tmp = tempfile.mkdtemp(prefix="evc_bench_")
    path = os.path.join(tmp, "train%d.tfrecord" % rank)
    recs = []
    for i in range(B):
        recs.append(R.make_sequence_example(
            f"v{i}", sorted(rng.choice(4716, size=3, replace=False).tolist()),
            {"rgb": rng.integers(0, 256, size=(300, 1024), dtype=np.uint8),
             "audio": rng.integers(0, 256, size=(300, 128), dtype=np.uint8)}))
    R.write_tfrecord(path, recs, with_crc=True)
    warm = 2
    rd = R.YT8MFrameFeatureReader(feature_names=["rgb", "audio"], feature_sizes=[1024, 128])
    threads = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
    it = rd.batches([path] * (steps + warm), B, native=True, verify_crc=True, prefetch=2, num_threads=threads)
    xq = torch.empty(B, 300, 1152, dtype=torch.uint8, device=dev)
    nfd = torch.empty(B, dtype=torch.int32, device=dev)
    lab = torch.empty(B, 4716, dtype=torch.bool, device=dev)
    out_host = torch.empty(tr.losses.numel(), dtype=torch.float32).pin_memory()
    topk_host = torch.empty(B, 20, dtype=torch.int32).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = 0.0
    for i, (ids, x, y, nf) in enumerate(it):
        if i == warm:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
        xq.copy_(x, non_blocking=True)
        nfd.copy_(nf, non_blocking=True)
        lab.copy_(y, non_blocking=True)
        tr.step(xq, nfd, lab)
        idx, _, _ = ops.topk(tr.s_eng.pred, 20)
        out_host.copy_(tr.losses, non_blocking=True)
        topk_host.copy_(idx, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    try:
        os.remove(path)
        os.rmdir(tmp)
    except OSError:
        pass
    return {"ms_per_step": ms, "wall_ms_per_step": wall_ms, "shard_bytes_per_step": B * 300 * 1152,
            "reader": "libevc_reader, CRC32C verified, prefetch 2", "reader_threads": threads}


def _profile_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    --set full summary of this round (profiles/r02_ncu_fwd_gemm.json written by scripts/ncu_summary.py); the
    bench itself never runs under a profiler."""
    for name in ("r02_ncu_fwd_gemm.json", "r01_ncu_fwd_gemm.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                with open(p) as f:
                    d = json.load(f)
                return float(d["dram_bytes_per_launch"]), "profiles/" + name
            except Exception:   # noqa: BLE001
                pass
    return 100.8e6, "profiles/r01_ncu_kernels_summary.txt (64.5 MB read + 36.3 MB written per launch)"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from efficientvideoclassification_youtube8m_b200 import _lib, ops
    from efficientvideoclassification_youtube8m_b200 import synthetic as O   # input generator (no oracle here)
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import (StudentEvaluator, StudentFinetuneTrainer,
                                                                   TeacherEvaluator, TeacherStudentTrainer)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # (the image's default prints a version banner on stdout)
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    finetune = args.workload == "finetune_cfg4"
    cfg = ModelConfig(lstm_cells=2048, num_mixtures=4) if finetune else ModelConfig()
    # Reference hyper-parameters (run_train.sh / train.py defaults: Adam, clip 1.0, regularization_penalty 2) except
    # the learning rate.  On THESE synthetic inputs (uniform random features, random labels) the reference graph
    # itself diverges at its default lr 1e-3: the float32 CPU restatement reaches |state| = 18 and a NaN L_PRED at
    # step 4, rotating batches or not (tests/divergence_full_size.py, recorded in
    # profiles/r02_cpu_reference_divergence_lr1e-3.json; check_numerics would abort the reference there), while it
    # stays finite at 1e-5 (profiles/r02_cpu_reference_finite_lr1e-5.json).  The optimizer does identical work at any lr; the
    # step sees NB different batches in rotation, like a training run.
    if finetune:   # BASELINE configs[3]: student fine-tune, lstm_cells 2048, 4 mixtures, clip 1.0
        tr = StudentFinetuneTrainer(cfg, batch_size=B, device=dev, base_learning_rate=args.lr, precise=args.precise)
    else:
        tr = TeacherStudentTrainer(cfg, batch_size=B, device=dev, base_learning_rate=args.lr, precise=args.precise)

    # synthetic batches: NB quantised (uint8, what a tfrecord holds) host-pinned batches; the device-resident
    # timing uses their dequantised float32 form (the tensor the reference's reader hands to the graph)
    NB = args.batches
    hq, hlab, dxf = [], [], []
    for i in range(NB):
        rng = np.random.default_rng(1234 + 1000 * rank + i)
        q = rng.integers(0, 256, size=(B, 300, 1152), dtype=np.uint8)
        _, _, lab = O.synthetic_batch(B, seed=4321 + 1000 * rank + i, num_features=4, full_length=True)
        hq.append(torch.from_numpy(q).pin_memory())
        hlab.append(torch.from_numpy(lab).view(torch.uint8).pin_memory())
    host_nf = torch.full((B,), 300, dtype=torch.int32).pin_memory()
    dnf = host_nf.to(dev)
    dlab = [h.to(dev) for h in hlab]
    for i in range(NB):                    # Dequantize on the device (bit-identical to utils.Dequantize, tests)
        xf = torch.empty(B, 300, 1152, dtype=torch.float32, device=dev)
        ops.frames_pack_u8(hq[i].to(dev), dnf, None, 300, 1, False, out_f32=xf)
        dxf.append(xf)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps

    # ---------------- device-resident timing (value)
    # every step reads a different 354 MB batch and writes + re-reads >3 GB of activations: nothing a step needs
    # is left in the 126 MB L2 by the previous one
    sampler = ClockSampler(local_rank)
    state = {"i": 0}

    def resident_step():
        i = state["i"] % NB
        state["i"] += 1
        tr.step(dxf[i], dnf, dlab[i])

    n0 = _lib.launch_count()
    resident_step()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - n0
    sampler.start()
    ms_step = timed(resident_step, args.steps, args.warmup)
    try:
        losses = tr.fetch()
    except FloatingPointError as e:      # the reference's check_numerics would stop training here
        losses = {"error": str(e)}

    # ---------------- end to end through the public step API, HOST buffers (e2e): the quantised uint8 batch a
    # tfrecord holds travels from pinned host memory every step (double buffered on a copy stream), Dequantize
    # runs inside the pack kernel, the losses and the student's top-20 come back
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    out_host = torch.empty(tr.losses.numel(), dtype=torch.float32).pin_memory()
    topk_host = torch.empty(B, 20, dtype=torch.int32).pin_memory()
    pred_eng = tr.s_eng

    def make_e2e(host_x, dev_x):
        st = {"i": 0}
        dl = [torch.empty_like(hlab[0], device=dev) for _ in range(2)]
        dn = torch.empty_like(host_nf, device=dev)

        def prefetch(slot, i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[slot])
                dev_x[slot].copy_(host_x[i % NB], non_blocking=True)
                dn.copy_(host_nf, non_blocking=True)
                dl[slot].copy_(hlab[i % NB], non_blocking=True)
                ready[slot].record(copy_stream)

        def step():
            slot = st["i"] & 1
            prefetch(slot ^ 1, st["i"] + 1)          # next batch travels while this one computes
            main.wait_event(ready[slot])
            tr.step(dev_x[slot], dn, dl[slot])
            idx, val, _ = ops.topk(pred_eng.pred, 20)
            freed[slot].record(main)
            out_host.copy_(tr.losses, non_blocking=True)
            topk_host.copy_(idx, non_blocking=True)
            main.synchronize()                       # the step's result is on the host
            st["i"] += 1

        main.synchronize(); copy_stream.synchronize()
        freed[0].record(main); freed[1].record(main)
        prefetch(0, 0)
        return step

    e2e_steps = max(3, args.steps // 2)
    dq = [torch.empty_like(hq[0], device=dev) for _ in range(2)]
    ms_e2e = timed(make_e2e(hq, dq), e2e_steps, 2)
    h2d = hq[0].numel() + host_nf.numel() * 4 + hlab[0].numel()
    d2h = out_host.numel() * 4 + topk_host.numel() * 4
    del dq
    # the same loop fed with the dequantised float32 batch (4x the bytes over PCIe): what the reference's reader
    # queue hands to the graph
    e2e_f32 = None
    if not args.skip_f32_e2e:
        hf = [x.cpu().pin_memory() for x in dxf[:2]]
        df = [torch.empty_like(dxf[0]) for _ in range(2)]
        hx = [hf[i % 2] for i in range(NB)]
        ms_f32 = timed(make_e2e(hx, df), e2e_steps, 2)
        e2e_f32 = {"value": world * B / (ms_f32 * 1e-3), "unit": "videos/s", "ms_per_step": ms_f32,
                   "h2d_bytes_per_step": hf[0].numel() * 4 + host_nf.numel() * 4 + hlab[0].numel()}
        del hf, df, hx
    clocks = sampler.stop()          # sampled over the device-resident and e2e timed regions

    # ---------------- the whole input path: TFRecord shards on disk -> native reader (libevc_reader, uint8
    # batches decoded into pinned memory on a background thread) -> H2D -> step -> results on the host; every rank
    # reads its own shard.  Extra information only: a failure here must not cost the bench line.
    e2e_tfrecord = None
    if not finetune and not args.skip_tfrecord:
        try:
            r = _tfrecord_e2e(tr, B, dev, e2e_steps, ops, rank)
            ms_t = r["ms_per_step"]
            if world > 1:
                t = torch.tensor([ms_t], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_t = t.item()
            e2e_tfrecord = dict(r, value=world * B / (ms_t * 1e-3), unit="videos/s", ms_per_step=ms_t)
        except Exception as e:   # noqa: BLE001
            e2e_tfrecord = {"error": f"{type(e).__name__}: {e}"}
            if world > 1:
                raise

    # ---------------- dominant kernel alone: RNN_L1 cell-0 forward steps of the teacher (15 launches; the
    # student's 6 for the fine-tune workload)
    t = tr.s_eng if finetune else tr.t_eng
    pset = tr.student if finetune else tr.teacher
    H, D, R1, ell = cfg.lstm_cells, cfg.feature_size, t.R1, t.ell
    lay = t.l1[0]

    def l1_fwd():
        ops.lstm_seq_fwd(t.x, R1 * D, D, pset.shadow[pset.kernel(0, 0)],
                         pset.w[pset.bias(0, 0)], R1, H, ell, t.len_l1, lay.h_all, lay.c_all, lay.gates)
    ms_seq = timed(l1_fwd, 5, 3)
    flops_seq = 2.0 * R1 * 4 * H * (D * ell + H * (ell - 1))
    peaks, peak_kind = _peaks()
    achieved = flops_seq / (ms_seq * 1e-3) / 1e12
    traffic, traffic_src = _profile_traffic()
    roof = {"bound": "tensor", "kernel": "gemm_kernel<A=K,B=MN,BN=256,EPI_LSTM_FWD> (%s RNN_L1 cell 0, %d rows)"
            % ("student" if finetune else "teacher", R1),
            "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"],
            # measured by ncu --set full in a separate run (never in the bench): algorithmic bytes are 61 MB read
            # (x_t, h, W, c) + 73.5 MB written (c, h, gates) per launch, the L2 absorbs part of it
            "traffic": None if finetune else traffic, "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
            "peak_kind": peak_kind + " burst bf16", "launches_timed": ell, "avg_launch_ms": ms_seq / ell}

    # ---------------- BASELINE configs #2, #4, #5 (one GPU, device resident, a few steps each)
    infer, configs = None, None
    if world == 1 and not args.skip_infer and not finetune:
        Bi = 1024
        xi, nfi, _ = O.synthetic_batch(Bi, seed=99, full_length=True)
        dxi, dnfi = torch.from_numpy(xi).to(dev), torch.from_numpy(nfi).to(dev)
        ev = StudentEvaluator(tr.student, Bi)
        ms_inf = timed(lambda: ev.step(dxi, dnfi), 10, 3)
        infer = {"student_infer_videos_per_s": Bi / (ms_inf * 1e-3), "batch": Bi, "ms_per_step": ms_inf}
        del ev
        # the teacher on all 300 frames (validate.py:149-155): denominator of the paper's inference-cost ratio
        evt = TeacherEvaluator(tr.teacher, Bi)
        ms_t = timed(lambda: evt.step(dxi, dnfi), 5, 2)
        infer.update({"teacher_infer_videos_per_s": Bi / (ms_t * 1e-3), "teacher_ms_per_step": ms_t,
                      "student_speedup_over_teacher": ms_t / ms_inf})
        del evt
        if not args.skip_configs:
            configs = {}
            # #5: frame-budget sweep and random vs uniform sampling (student fine-tune step B=256, student inference
            # B=1024; the student's own weights, cfg #1 dimensions)
            sweep = {}
            for every_n, sampling in ((5, "uniform"), (10, "uniform"), (20, "uniform"), (30, "uniform"),
                                      (10, "random_frames"), (10, "random_sequence")):
                ft = StudentFinetuneTrainer(ModelConfig(), batch_size=B, device=dev, every_n=every_n,
                                            base_learning_rate=args.lr, sampling=sampling)
                ms_ft = timed(lambda: ft.step(dxf[0], dnf, dlab[0]), 5, 3)
                evs = StudentEvaluator(ft.student, Bi, every_n=every_n, sampling=sampling)
                ms_ev = timed(lambda: evs.step(dxi, dnfi), 5, 3)
                sweep[f"every_n={every_n},{sampling}"] = {
                    "student_frames": ft.student_frames, "train_videos_per_s": B / (ms_ft * 1e-3),
                    "train_ms_per_step": ms_ft, "infer_videos_per_s": Bi / (ms_ev * 1e-3), "infer_ms_per_step": ms_ev}
                del ft, evs
            configs["cfg5_frame_budget_and_sampling"] = sweep
            # #4: student fine-tune with 2048 cells, 4 mixtures, clip 1.0
            ft4 = StudentFinetuneTrainer(ModelConfig(lstm_cells=2048, num_mixtures=4), batch_size=B, device=dev,
                                         base_learning_rate=args.lr)
            ms4 = timed(lambda: ft4.step(dxf[0], dnf, dlab[0]), 5, 3)
            configs["cfg4_student_finetune_2048x4"] = {
                "train_videos_per_s": B / (ms4 * 1e-3), "ms_per_step": ms4,
                "model_tflops": B / (ms4 * 1e-3) * GF_PER_VIDEO_FINETUNE_CFG4 / 1e3}
            del ft4
            # the joint step in split-bf16 "precise" mode (3 tensor-core products per contraction)
            if not args.precise:
                trp = TeacherStudentTrainer(ModelConfig(), batch_size=B, device=dev, base_learning_rate=args.lr,
                                            precise=True)
                msp = timed(lambda: trp.step(dxf[0], dnf, dlab[0]), 3, 2)
                configs["cfg1_precise_split_bf16"] = {"train_videos_per_s": B / (msp * 1e-3), "ms_per_step": msp}
                del trp
        del dxi

    if rank == 0:
        vps = world * B / (ms_step * 1e-3)
        e2e_vps = world * B / (ms_e2e * 1e-3)
        gf = GF_PER_VIDEO_FINETUNE_CFG4 if finetune else GF_PER_VIDEO_TRAIN
        line = {
            "metric": "H-LSTM student fine-tune train videos/s" if finetune else
                      "H-LSTM teacher-student train videos/s", "value": vps, "unit": "videos/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16x2" if args.precise else "bf16", "data": "synthetic",
            "config": {"workload": "student fine-tune step (run_finetune.sh with lstm_cells 2048, 4 mixtures; "
                                   "BASELINE configs[3])" if finetune else
                                   "teacher-student joint train step (run_train.sh defaults; BASELINE configs[0]/[2])",
                       "batch_per_gpu": B, "global_batch": world * B, "frames": 300, "features": 1152,
                       "every_n": 10, "lstm_cells": cfg.lstm_cells, "lstm_layers": 2,
                       "moe_num_mixtures": cfg.num_mixtures,
                       "classes": 4716, "parallelism": f"dp{world}", "base_learning_rate": args.lr,
                       "rotating_batches": NB,
                       "l2_policy": "a different 354 MB batch every step + >3 GB of activations per step exceed the "
                                    "126 MB L2"},
            "model_tflops": vps * gf / 1e3,
            "frac_of_sustained_bf16_peak": vps * gf / 1e3 / world / peaks["bf16_tflops_sustained"],
            "e2e": {"value": e2e_vps, "unit": "videos/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e, "input": "uint8 features as stored in the tfrecords, pinned host memory"},
            "e2e_f32_input": e2e_f32,
            "e2e_tfrecord_input": e2e_tfrecord,
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks, "roofline": roof, "losses": losses,
        }
        if infer:
            line["student_infer"] = infer
        if configs:
            line["configs"] = configs
        if world == 1 and not args.skip_cpu and not finetune:
            v, sec, cores = cpu_baseline(32, 2, 1)
            line["cpu_baseline"] = {"value": v, "unit": "videos/s", "cores": cores, "kind": "port",
                                    "sample": "2 timed steps (after 1 warm-up) of a 32-video batch of the same T+S "
                                              "train step, f32 PyTorch-CPU restatement, all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--lr", type=float, default=1e-5,
                    help="base_learning_rate (train.py:75 default 1e-3 diverges on random synthetic inputs, also in "
                         "the CPU restatement: profiles/r02_cpu_reference_divergence_lr1e-3.json)")
    ap.add_argument("--batches", type=int, default=8, help="distinct synthetic batches in rotation")
    ap.add_argument("--precise", action="store_true", help="split-bf16 mode (3 products per contraction)")
    ap.add_argument("--skip-f32-e2e", action="store_true")
    ap.add_argument("--skip-configs", action="store_true")
    ap.add_argument("--workload", default="ts_train", choices=["ts_train", "finetune_cfg4"])
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-infer", action="store_true")
    ap.add_argument("--skip-tfrecord", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
