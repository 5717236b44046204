"""Throughput of the input path (videos/s): pure-Python decoder vs libevc_reader at several thread counts, on
synthetic YT8M-shaped shards (300 frames x (1024 rgb + 128 audio) uint8, ~346 KB per video).  CPU only.

    python scripts/bench_reader.py [videos_per_shard] [shards]
"""
import ctypes
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from efficientvideoclassification_youtube8m_b200 import readers as R

per_shard = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shards = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(0)
d = tempfile.mkdtemp()
paths = []
for s in range(shards):
    recs = []
    for i in range(per_shard):
        n = int(rng.integers(100, 301))
        recs.append(R.make_sequence_example(f"v{s}_{i}", sorted(rng.choice(4716, size=3, replace=False).tolist()),
                                            {"rgb": rng.integers(0, 256, size=(n, 1024), dtype=np.uint8),
                                             "audio": rng.integers(0, 256, size=(n, 128), dtype=np.uint8)}))
    paths.append(os.path.join(d, f"train{s}.tfrecord"))
    R.write_tfrecord(paths[-1], recs, with_crc=True)
mb = sum(os.path.getsize(p) for p in paths) / 1e6
print(f"{shards} shards, {shards * per_shard} videos, {mb:.0f} MB; host threads: {os.cpu_count()}")
rd = R.YT8MFrameFeatureReader(feature_names=["rgb", "audio"], feature_sizes=[1024, 128])

t = time.time()
n = sum(len(b[0]) for b in rd.batches(paths[:1], 64, pin_memory=False, native=False))
print(f"python decoder            {n / (time.time() - t):9.0f} videos/s")

lib = R.reader_lib()
B = 256
x = torch.zeros(B, 300, 1152, dtype=torch.uint8)
y = torch.zeros(B, 4716, dtype=torch.uint8)
nf = torch.zeros(B, dtype=torch.int32)
ids = ctypes.create_string_buffer(B * 64)
cp = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])
names = (ctypes.c_char_p * 2)(b"rgb", b"audio")
sizes = (ctypes.c_int * 2)(1024, 128)
for crc in (0, 1):
    for threads in (1, 2, 4, 8, 16, 32):
        if threads > 2 * (os.cpu_count() or 1):
            break
        h = lib.evc_reader_open(cp, len(paths), names, sizes, 2, 4716, 300, threads, crc)
        tot, t = 0, time.time()
        for _ in range(4):
            lib.evc_reader_rewind(h)
            while True:
                g = lib.evc_reader_next(h, B, x.data_ptr(), y.data_ptr(), nf.data_ptr(), ids, 64)
                assert g >= 0, lib.evc_reader_last_error()
                if g == 0:
                    break
                tot += g
        dt = time.time() - t
        lib.evc_reader_close(h)
        print(f"native, {threads:2d} threads, crc={crc} {tot / dt:9.0f} videos/s  ({tot / dt * 0.3456 / 1e3:.2f} GB/s of uint8 batches)")
