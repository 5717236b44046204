"""Synthetic YouTube-8M-shaped inputs for benchmarks and examples: dequantised-uint8
frame features (utils.py:21-25 Dequantize with the readers.py:178-179 defaults max=2, min=-2),
zero-padded past num_frames (readers.py:173), a few labels per video.  Pure numpy, host side."""
from __future__ import annotations

import numpy as np

MAX_FRAMES = 300


def synthetic_batch(batch, seed=1234, num_features=1152, vocab_size=4716, full_length=False,
                    max_frames=MAX_FRAMES):
    """Returns (features f32 [B,300,D], num_frames int32 [B], labels bool [B,V])."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, size=(batch, max_frames, num_features), dtype=np.uint8)
    x = q.astype(np.float32) * np.float32(4.0 / 255.0) + np.float32(4.0 / 512.0 - 2.0)
    if full_length:
        nf = np.full((batch,), max_frames, dtype=np.int32)
    else:
        nf = rng.integers(1, max_frames + 1, size=(batch,)).astype(np.int32)
    x[np.arange(max_frames)[None, :] >= nf[:, None]] = 0.0
    lrng = np.random.default_rng(seed + 3087)
    labels = np.zeros((batch, vocab_size), dtype=bool)
    for b in range(batch):
        k = min(max(1, int(lrng.poisson(3.4))), vocab_size)
        labels[b, lrng.choice(vocab_size, size=k, replace=False)] = True
    return x, nf, labels
