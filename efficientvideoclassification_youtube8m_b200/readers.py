"""Frame-level YouTube-8M input path (code_student_uniform/readers.py:114-246, utils.py:10-25) without
TensorFlow: TFRecord framing and the SequenceExample protobuf are decoded by hand on the host, and
the features stay **uint8** — Dequantize, the zero padding to max_frames and the l2-normalise run
fused on the GPU (`evc_frames_pack_u8`), so 4x fewer bytes cross PCIe than with the reference's
float32 batches.

    reader = YT8MFrameFeatureReader(feature_names=["rgb", "audio"], feature_sizes=[1024, 128])
    for ids, feats_u8, labels, num_frames in reader.batches(glob("train*.tfrecord"), 256):
        trainer.step(feats_u8.cuda(non_blocking=True), num_frames.cuda(), labels.cuda())
"""
from __future__ import annotations

import ctypes
import os
import queue
import struct
import threading
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

# ------------------------------------------------------------------ native decoder (include/evc_reader.h)
READER_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libevc_reader.so")
_reader_lib = None


def reader_lib():
    """ctypes binding of libevc_reader.so (csrc/evc_reader.cpp: multi-threaded TFRecord + SequenceExample
    decoder writing straight into the caller's pinned batch buffers)."""
    global _reader_lib
    if _reader_lib is None:
        if not os.path.exists(READER_LIB_PATH):
            raise ImportError(f"{READER_LIB_PATH} not found: build it with "
                              "`python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
        lib = ctypes.CDLL(READER_LIB_PATH)
        P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        lib.evc_reader_version.argtypes, lib.evc_reader_version.restype = [], I
        lib.evc_reader_last_error.argtypes, lib.evc_reader_last_error.restype = [], ctypes.c_char_p
        lib.evc_reader_open.argtypes = [ctypes.POINTER(ctypes.c_char_p), I, ctypes.POINTER(ctypes.c_char_p),
                                        ctypes.POINTER(I), I, I, I, I, I]
        lib.evc_reader_open.restype = P
        lib.evc_reader_next.argtypes, lib.evc_reader_next.restype = [P, I, P, P, P, P, I], I
        lib.evc_reader_position.argtypes, lib.evc_reader_position.restype = [P], L
        lib.evc_reader_rewind.argtypes, lib.evc_reader_rewind.restype = [P], I
        lib.evc_reader_close.argtypes, lib.evc_reader_close.restype = [P], None
        lib.evc_crc32c_masked.argtypes, lib.evc_crc32c_masked.restype = [ctypes.c_char_p, L], ctypes.c_uint
        _reader_lib = lib
    return _reader_lib


class BaseReader(object):
    """Inherit from this class when implementing new readers."""

    def prepare_reader(self, unused_filename_queue):
        raise NotImplementedError()


# ------------------------------------------------------------------ protobuf wire format (just enough)
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yields (field number, wire type, value) of one message; value = int (varint) or bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _parse_feature(buf: bytes):
    """tensorflow.Feature -> ('bytes', [bytes...]) | ('int64', [int...]) | ('float', ndarray)."""
    for num, _, v in _fields(buf):
        if num == 1:      # BytesList
            return "bytes", [x for n, _, x in _fields(v) if n == 1]
        if num == 3:      # Int64List (packed or not)
            out: List[int] = []
            for n, wt, x in _fields(v):
                if n != 1:
                    continue
                if wt == 0:
                    out.append(x)
                else:
                    p = 0
                    while p < len(x):
                        val, p = _varint(x, p)
                        out.append(val)
            return "int64", [o - (1 << 64) if o >= (1 << 63) else o for o in out]
        if num == 2:      # FloatList
            chunks = [x for n, _, x in _fields(v) if n == 1]
            return "float", np.frombuffer(b"".join(chunks), dtype="<f4")
    return "empty", []


def _parse_map(buf: bytes, value_parser) -> Dict[str, object]:
    """map<string, X> entries (field 1 = key, field 2 = value) of a Features / FeatureLists message."""
    out = {}
    for num, _, entry in _fields(buf):
        if num != 1:
            continue
        key, val = None, None
        for n, _, x in _fields(entry):
            if n == 1:
                key = x.decode("utf-8")
            elif n == 2:
                val = value_parser(x)
        out[key] = val
    return out


def parse_sequence_example(buf: bytes):
    """tensorflow.SequenceExample -> (context {name: feature}, feature_lists {name: [feature, ...]})."""
    context, lists = {}, {}
    for num, _, v in _fields(buf):
        if num == 1:
            context = _parse_map(v, _parse_feature)
        elif num == 2:
            lists = _parse_map(v, lambda fl: [_parse_feature(x) for n, _, x in _fields(fl) if n == 1])
    return context, lists


def tfrecord_iterator(path: str) -> Iterator[bytes]:
    """TFRecord framing: uint64 length, uint32 crc, payload, uint32 crc (CRCs are not verified)."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if len(head) < 12:
                return
            (length,) = struct.unpack("<Q", head[:8])
            data = f.read(length)
            if len(data) < length:
                raise IOError(f"truncated record in {path}")
            f.read(4)
            yield data


# ------------------------------------------------------------------ writer (tests / synthetic shards)
def _enc_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(num: int, payload: bytes) -> bytes:
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def make_sequence_example(video_id: str, labels: Sequence[int], features: Dict[str, np.ndarray]) -> bytes:
    """features: {name: uint8 [num_frames, size]} -> serialized SequenceExample in the YT8M layout."""
    def bytes_feature(b):
        return _ld(1, _ld(1, b))

    def int64_feature(vals):
        return _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in vals)))

    ctx = _ld(1, _ld(1, b"id") + _ld(2, bytes_feature(video_id.encode()))) + \
        _ld(1, _ld(1, b"labels") + _ld(2, int64_feature(labels)))
    fl = b""
    for name, mat in features.items():
        flist = b"".join(_ld(1, bytes_feature(np.ascontiguousarray(row, dtype=np.uint8).tobytes())) for row in mat)
        fl += _ld(1, _ld(1, name.encode()) + _ld(2, flist))
    return _ld(1, ctx) + _ld(2, fl)


def write_tfrecord(path: str, records: Iterable[bytes], with_crc: bool = False) -> None:
    """TFRecord framing; with_crc=True stores the masked CRC32C of the length and of the payload like
    TensorFlow's writer (computed by the native library), otherwise zeros."""
    crc = reader_lib().evc_crc32c_masked if with_crc else None
    with open(path, "wb") as f:
        for r in records:
            head = struct.pack("<Q", len(r))
            c1 = struct.pack("<I", crc(head, 8)) if crc else b"\0\0\0\0"
            c2 = struct.pack("<I", crc(r, len(r))) if crc else b"\0\0\0\0"
            f.write(head + c1 + r + c2)


# ------------------------------------------------------------------ the reader
class YT8MFrameFeatureReader(BaseReader):
    """Reads TFRecords of SequenceExamples with a sparse int64 'labels' context feature and one
    byte-quantised feature list per name in `feature_names` (readers.py:114-246)."""

    def __init__(self, num_classes=4716, feature_sizes=(1024,), feature_names=("inc3",), max_frames=300):
        assert len(feature_names) == len(feature_sizes), \
            "length of feature_names (={}) != length of feature_sizes (={})".format(len(feature_names), len(feature_sizes))
        assert len(feature_names) > 0, "No feature selected: feature_names is empty!"
        self.num_classes = num_classes
        self.feature_sizes = list(feature_sizes)
        self.feature_names = list(feature_names)
        self.max_frames = max_frames

    def prepare_reader(self, filenames):
        """Yields (video_id str, quantised features uint8 [max_frames, sum(sizes)] zero-filled past
        num_frames, labels bool [num_classes], num_frames int) per video."""
        D = sum(self.feature_sizes)
        for path in filenames:
            for rec in tfrecord_iterator(path):
                ctx, lists = parse_sequence_example(rec)
                vid = ctx["id"][1][0].decode("utf-8")
                labels = np.zeros(self.num_classes, dtype=bool)
                idx = np.asarray(ctx["labels"][1], dtype=np.int64)
                labels[idx[(idx >= 0) & (idx < self.num_classes)]] = True     # sparse_to_dense(validate_indices=False)
                mat = np.zeros((self.max_frames, D), dtype=np.uint8)
                num_frames, col = -1, 0
                for name, size in zip(self.feature_names, self.feature_sizes):
                    rows = lists[name]
                    n = len(rows)
                    if num_frames == -1:
                        num_frames = n
                    elif n != num_frames:
                        raise ValueError(f"{vid}: feature '{name}' has {n} frames, expected {num_frames}")
                    k = min(n, self.max_frames)
                    if k:
                        block = np.frombuffer(b"".join(r[1][0] for r in rows[:k]), dtype=np.uint8).reshape(k, size)
                        mat[:k, col:col + size] = block
                    col += size
                yield vid, mat, labels, min(max(num_frames, 0), self.max_frames)

    def batches(self, filenames, batch_size, drop_remainder=False, pin_memory=None, native=True,
                num_threads=0, verify_crc=False, prefetch=0):
        """(ids, uint8 [B,max_frames,D], bool [B,num_classes], int32 [B]) torch tensors (pinned when CUDA is
        available) ready for `Trainer.step`.

        native=True: the shards are decoded by libevc_reader (num_threads worker threads, 0 = one per hardware
        thread) straight into the pinned batch tensors; prefetch=N decodes up to N batches ahead on a
        background thread (the C call releases the GIL), so decoding overlaps the GPU step.
        native=False: the pure-Python decoder above (reference implementation of the wire format, pinned
        against the protobuf library in tests/test_readers.py)."""
        pin = torch.cuda.is_available() if pin_memory is None else pin_memory
        if native:
            # prefetch > 0: a ring of prefetch + 3 pinned buffer sets is reused: one being filled by the worker,
            # `prefetch` queued, one in the consumer's hands and one more that the consumer handed to the GPU
            # in the previous iteration.  LIFETIME CONTRACT: the tensors of batch j stay untouched until the
            # consumer asks for batch j+2, i.e. an asynchronous `x.cuda(non_blocking=True)` of batch j must have
            # completed by the time the consumer requests batch j+2 (true for any loop that fetches a result of
            # step j, or synchronises the stream, once per iteration; otherwise record an event after the copy and
            # wait for it before calling next()).  prefetch = 0 allocates fresh tensors for every batch.
            it = self._native_batches(list(filenames), batch_size, drop_remainder, pin, num_threads, verify_crc,
                                      ring=prefetch + 3 if prefetch > 0 else 0)
            return _prefetched(it, prefetch) if prefetch > 0 else it
        return self._python_batches(filenames, batch_size, drop_remainder, pin)

    def _native_batches(self, filenames, batch_size, drop_remainder, pin, num_threads, verify_crc, ring=0):
        lib = reader_lib()
        D, V, T = sum(self.feature_sizes), self.num_classes, self.max_frames
        paths = (ctypes.c_char_p * len(filenames))(*[os.fsencode(p) for p in filenames])
        names = (ctypes.c_char_p * len(self.feature_names))(*[n.encode() for n in self.feature_names])
        sizes = (ctypes.c_int * len(self.feature_sizes))(*self.feature_sizes)
        handle = lib.evc_reader_open(paths, len(filenames), names, sizes, len(self.feature_names), V, T,
                                     int(num_threads), int(bool(verify_crc)))
        if not handle:
            raise IOError(lib.evc_reader_last_error().decode())
        id_stride = 64

        def alloc():
            return (torch.empty(batch_size, T, D, dtype=torch.uint8, pin_memory=pin),
                    torch.empty(batch_size, V, dtype=torch.uint8, pin_memory=pin),
                    torch.empty(batch_size, dtype=torch.int32, pin_memory=pin))

        buffers = [alloc() for _ in range(ring)]
        turn = 0
        try:
            while True:
                x, y, n = buffers[turn % ring] if ring else alloc()
                turn += 1
                ids = ctypes.create_string_buffer(batch_size * id_stride)
                got = lib.evc_reader_next(handle, batch_size, x.data_ptr(), y.data_ptr(), n.data_ptr(), ids, id_stride)
                if got < 0:
                    raise ValueError(lib.evc_reader_last_error().decode())
                if got == 0 or (got < batch_size and drop_remainder):
                    return
                raw = ids.raw
                vid = [raw[i * id_stride:(i + 1) * id_stride].split(b"\0", 1)[0].decode("utf-8") for i in range(got)]
                yield vid, x[:got], y[:got].view(torch.bool), n[:got]
                if got < batch_size:
                    return
        finally:
            lib.evc_reader_close(handle)

    def _python_batches(self, filenames, batch_size, drop_remainder, pin):
        ids, feats, labs, nfs = [], [], [], []

        def emit():
            x = torch.from_numpy(np.stack(feats))
            y = torch.from_numpy(np.stack(labs))
            n = torch.tensor(nfs, dtype=torch.int32)
            if pin:
                x, y, n = x.pin_memory(), y.pin_memory(), n.pin_memory()
            return list(ids), x, y, n

        for vid, mat, labels, nf in self.prepare_reader(filenames):
            ids.append(vid); feats.append(mat); labs.append(labels); nfs.append(nf)
            if len(ids) == batch_size:
                yield emit()
                ids, feats, labs, nfs = [], [], [], []
        if ids and not drop_remainder:
            yield emit()


def get_input_data_batches(reader: "YT8MFrameFeatureReader", data_pattern: str, batch_size: int = 1000,
                           num_epochs: Optional[int] = None, shuffle: bool = True, seed: int = 0,
                           rank: int = 0, world: int = 1, uneven: str = "min", agree=None, **batch_kwargs):
    """train.py:125-175 `get_input_data_tensors` as a generator of host batches: glob the pattern (IOError with
    the reference's message when nothing matches), shuffle the FILE order every epoch
    (`string_input_producer(files, num_epochs, shuffle=True)`), run `num_epochs` passes (None = forever) and
    yield `reader.batches(...)` tuples; the last batch of an epoch may be smaller
    (`allow_smaller_final_batch=True`).  The example-level mixing of `shuffle_batch_join` (a random queue of
    50 batches) is not reproduced: videos keep their order inside a shard.

    rank/world: data-parallel ranks read disjoint slices `order[rank::world]` of the (shuffled) file list, so
    every shard is read by exactly one rank.  YT8M shards hold different numbers of videos, so the ranks would
    produce different numbers of batches per epoch and the first one to finish would leave the others hanging
    in the step's collectives.  Before every batch the ranks therefore agree (one tiny all-reduce through
    `agree`, default: torch.distributed on the default group) on whether the epoch goes on:
      uneven="min"  the epoch ends for everybody when the first rank runs out (training: the few surplus
                    videos of the longer ranks are skipped this epoch, the shuffle moves them next epoch);
      uneven="pad"  the epoch ends when the LAST rank runs out; ranks that ran out yield empty batches
                    (0 videos, right shapes) so that no video is lost (evaluation: `steps.evaluation_loop`
                    pads them to the plan's batch size and drops the padding before the metrics)."""
    import glob
    files = sorted(glob.glob(data_pattern))
    if not files:
        raise IOError("Unable to find training files. data_pattern='" + data_pattern + "'.")
    if uneven not in ("min", "pad"):
        raise ValueError("uneven must be 'min' or 'pad'")
    if world > 1 and agree is None:
        agree = _dist_agree
    D, V, T = sum(reader.feature_sizes), reader.num_classes, reader.max_frames
    epoch = 0
    while num_epochs is None or epoch < num_epochs:
        order = list(files)
        if shuffle:
            np.random.default_rng(seed + epoch).shuffle(order)      # same permutation on every rank
        if world > 1:
            order = order[rank::world]
        it = iter(reader.batches(order, batch_size, **batch_kwargs)) if order else iter(())
        while True:
            batch = next(it, None)
            if world > 1:
                # have = 1 if this rank still has a batch; min over ranks = everybody has one, max = somebody has
                everybody, somebody = agree(batch is not None)
                if uneven == "min" and not everybody:
                    break
                if uneven == "pad":
                    if not somebody:
                        break
                    if batch is None:
                        batch = ([], torch.empty(0, T, D, dtype=torch.uint8), torch.empty(0, V, dtype=torch.bool),
                                 torch.empty(0, dtype=torch.int32))
            elif batch is None:
                break
            yield batch
        epoch += 1


def _dist_agree(have: bool):
    """(all ranks have a batch, some rank has a batch) over the default process group."""
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([1 if have else 0, 0 if have else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    have_n, miss_n = t.tolist()
    return miss_n == 0, have_n > 0


def _prefetched(it, depth):
    """Runs the iterator `it` on a background thread, `depth` items ahead."""
    q: "queue.Queue" = queue.Queue(maxsize=depth)
    done = object()

    def worker():
        try:
            for item in it:
                q.put(item)
            q.put(done)
        except BaseException as e:      # re-raised in the consumer
            q.put(e)

    threading.Thread(target=worker, daemon=True).start()
    while True:
        item = q.get()
        if item is done:
            return
        if isinstance(item, BaseException):
            raise item
        yield item
