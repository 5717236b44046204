"""TensorFlow checkpoint files ("tensor bundle", the V2 format `tf.train.Saver` writes since TF 1.0) without
TensorFlow: what `saver.restore(sess, tf.train.latest_checkpoint(train_dir))` reads in validate.py:216-224,
eval_finetune.py:202-209, train.py:592-598 and what train_convert_model.py:496-517 re-saves.

A checkpoint `<prefix>` is
  <prefix>.index                 a leveldb-format sorted string table [TF core/lib/io/table*, format.cc]:
                                 key ""    -> BundleHeaderProto {num_shards=1, endianness=2, version=3}
                                 key name  -> BundleEntryProto  {dtype=1, shape=2, shard_id=3, offset=4, size=5,
                                                                 crc32c=6 (fixed32, masked), slices=7}
                                 [TF core/protobuf/tensor_bundle.proto, core/util/tensor_bundle/tensor_bundle.cc]
  <prefix>.data-0000i-of-0000n   the raw little-endian tensor bytes of shard i, at [offset, offset+size)
  <train_dir>/checkpoint         CheckpointState text proto: model_checkpoint_path: "<prefix basename>"

Table format: data blocks of prefix-compressed entries (varint shared, non_shared, value_len, key delta, value)
followed by a restart array (fixed32 offsets + fixed32 count), each block followed by a 5-byte trailer (compression
type, masked crc32c of block+type); then the metaindex block, the index block (last key of a data block ->
BlockHandle varint offset, varint size) and the 48-byte footer (metaindex handle, index handle, padding, magic
0xdb4775248b80fb57).  Blocks are read uncompressed or snappy-compressed; the writer emits uncompressed blocks, which
is what TF's BundleWriter does.

PARITY NOTE: restated from the published format; no TensorFlow exists in this image to produce a reference file, so
the reader is pinned against the writer below and against hand-assembled blocks with prefix compression, several
restart intervals, snappy blocks and sharded data files (tests/test_tf_checkpoint.py).
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# DataType enum values of tensorflow/core/framework/types.proto that a Saver checkpoint of this model family holds
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
           6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("bool"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
           22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DTYPE_ENUM = {v: k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------ crc32c (Castagnoli), masked as TF stores it
def _make_crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return np.array(tab, dtype=np.uint32)


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes) -> int:
    """Software CRC32C; large buffers go through the native reader library when it is built."""
    if len(data) >= 4096:
        try:
            from .readers import reader_lib
            return _unmask(int(reader_lib().evc_crc32c_masked(data, len(data))))
        except Exception:       # noqa: BLE001 - library not built: fall through to the table loop
            pass
    c = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def _unmask(masked: int) -> int:
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / minimal protobuf
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("corrupted varint")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _pb_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """(field number, wire type, value) of one protobuf message (value: int for varint/fixed, bytes for length-delimited)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for num, _, v in _pb_fields(buf):
        if num == 2:                                   # repeated Dim dim
            size = 0
            for n2, _, v2 in _pb_fields(v):
                if n2 == 1:
                    size = v2 - (1 << 64) if v2 >> 63 else v2
            dims.append(size)
        elif num == 3 and v:                           # unknown_rank
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for num, _, v in _pb_fields(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            e["shape"] = _parse_shape(v)
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = v
        elif num == 7:
            e["slices"] += 1
    return e


def _entry_proto(dtype_enum: int, shape: Tuple[int, ...], shard: int, offset: int, size: int, crc: int) -> bytes:
    shp = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype_enum) + b"\x12" + _put_varint(len(shp)) + shp
    if shard:
        out += b"\x18" + _put_varint(shard)
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ------------------------------------------------------------------ snappy (raw format) decompression
def snappy_decompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupted snappy block")
        for _ in range(ln):                            # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupted snappy block (length)")
    return bytes(out)


# ------------------------------------------------------------------ table reader
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + 5 > len(data):
        raise ValueError("block handle past the end of the table")
    block, ctype = data[offset:offset + size], data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if _unmask(stored) != crc32c(data[offset:offset + size + 1]):
            raise ValueError("block checksum mismatch in checkpoint index")
    if ctype == 0:
        return block
    if ctype == 1:
        return snappy_decompress(block)
    raise ValueError(f"unknown block compression type {ctype}")


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise ValueError("corrupted block")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise ValueError("corrupted block entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(data: bytes, verify_checksums: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a leveldb-format table, in key order."""
    if len(data) < 48:
        raise ValueError("file too short to be an sstable")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError("not an sstable (bad magic number)")
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify_checksums)):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify_checksums)))
    return out


# ------------------------------------------------------------------ table writer
class _BlockBuilder:
    def __init__(self, restart_interval: int):
        self.interval, self.buf, self.restarts, self.count, self.last = restart_interval, bytearray(), [0], 0, b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def __len__(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def write_table(pairs: Iterable[Tuple[bytes, bytes]], block_size: int = 4096, restart_interval: int = 16) -> bytes:
    """Serialise sorted (key, value) pairs as a leveldb-format table (uncompressed blocks)."""
    out = bytearray()
    index = _BlockBuilder(1)

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    cur, last_key, prev = _BlockBuilder(restart_interval), None, None
    for key, value in pairs:
        if prev is not None and key <= prev:
            raise ValueError("table keys must be strictly increasing")
        prev = key
        cur.add(key, value)
        last_key = key
        if len(cur) >= block_size:
            index.add(last_key, emit(cur.finish()))
            cur = _BlockBuilder(restart_interval)
    if cur.count or last_key is None:
        index.add(last_key if last_key is not None else b"", emit(cur.finish()))
    meta = emit(_BlockBuilder(restart_interval).finish())
    idx = emit(index.finish())
    footer = (meta + idx).ljust(40, b"\x00") + struct.pack("<Q", TABLE_MAGIC)
    return bytes(out) + footer


# ------------------------------------------------------------------ tensor bundle
def _shard_name(prefix: str, shard: int, num_shards: int) -> str:
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def list_variables(prefix: str) -> Dict[str, Tuple[np.dtype, Tuple[int, ...]]]:
    """name -> (dtype, shape) of every tensor in the checkpoint (tf.train.list_variables)."""
    with open(prefix + ".index", "rb") as f:
        pairs = read_table(f.read())
    out = {}
    for k, v in pairs:
        if k == b"":
            continue
        e = _parse_entry(v)
        if e["dtype"] not in _DTYPES:
            continue                                   # strings / resources: not tensors of this model family
        out[k.decode("utf-8")] = (_DTYPES[e["dtype"]], e["shape"])
    return out


def load_variables(prefix: str, names: Optional[Iterable[str]] = None, verify_checksums: bool = True) -> Dict[str, np.ndarray]:
    """name -> array for `names` (default: every numeric tensor) of the checkpoint `<prefix>`."""
    with open(prefix + ".index", "rb") as f:
        pairs = read_table(f.read(), verify_checksums)
    header = dict(pairs).get(b"")
    if header is None:
        raise ValueError("checkpoint index has no bundle header")
    num_shards, endianness = 1, 0
    for num, _, v in _pb_fields(header):
        if num == 1:
            num_shards = v
        elif num == 2:
            endianness = v
    if endianness != 0:
        raise ValueError("big-endian checkpoints are not supported")
    want = None if names is None else set(names)
    entries = {}
    for k, v in pairs:
        name = k.decode("utf-8")
        if k == b"" or (want is not None and name not in want):
            continue
        entries[name] = _parse_entry(v)
    if want is not None and want - set(entries):
        raise KeyError(f"not in checkpoint {prefix}: {sorted(want - set(entries))}")
    files, out = {}, {}
    try:
        for name, e in entries.items():
            if e["dtype"] not in _DTYPES:
                if want is not None:
                    raise ValueError(f"{name}: unsupported dtype enum {e['dtype']}")
                continue
            if e["slices"]:
                raise ValueError(f"{name}: partitioned (sliced) variables are not supported")
            dt = _DTYPES[e["dtype"]]
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise ValueError(f"{name}: {e['size']} bytes do not match {dt} {e['shape']}")
            f = files.get(e["shard_id"])
            if f is None:
                f = files[e["shard_id"]] = open(_shard_name(prefix, e["shard_id"], num_shards), "rb")
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError(f"{name}: data shard truncated")
            if verify_checksums and e["crc32c"] is not None and _unmask(e["crc32c"]) != crc32c(raw):
                raise ValueError(f"{name}: tensor checksum mismatch")
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for f in files.values():
            f.close()
    return out


def save_variables(prefix: str, variables: Dict[str, np.ndarray]) -> None:
    """Write `<prefix>.index` + `<prefix>.data-00000-of-00001` holding name -> array (one shard, little endian,
    per-tensor masked crc32c), the layout `tf.train.Saver(write_version=V2).save` produces."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    pairs = [(b"", b"\x08\x01\x1a\x02\x08\x01")]       # num_shards = 1, version {producer = 1}
    offset = 0
    with open(_shard_name(prefix, 0, 1), "wb") as f:
        for name in sorted(variables, key=lambda s: s.encode("utf-8")):
            a = np.asarray(variables[name])               # (ascontiguousarray would turn scalars into 1-d)
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in _DTYPE_ENUM:
                raise ValueError(f"{name}: dtype {a.dtype} has no checkpoint encoding here")
            raw = a.astype(dt, copy=False).tobytes()
            f.write(raw)
            pairs.append((name.encode("utf-8"), _entry_proto(_DTYPE_ENUM[np.dtype(dt)], a.shape, 0, offset, len(raw),
                                                             _mask(crc32c(raw)))))
            offset += len(raw)
    with open(prefix + ".index", "wb") as f:
        f.write(write_table(pairs))


# ------------------------------------------------------------------ CheckpointState file (tf.train.latest_checkpoint)
def latest_checkpoint(train_dir: str) -> Optional[str]:
    """Prefix of the newest checkpoint recorded in `<train_dir>/checkpoint`, or None (tf.train.latest_checkpoint)."""
    path = os.path.join(train_dir, "checkpoint")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', f.read(), re.M)
    if not m:
        return None
    prefix = m.group(1)
    if not os.path.isabs(prefix):
        prefix = os.path.join(train_dir, prefix)
    return prefix if (os.path.exists(prefix + ".index") or os.path.exists(prefix + ".npz")) else None


def update_checkpoint_state(train_dir: str, prefix: str, keep: Optional[List[str]] = None) -> None:
    base = os.path.basename(prefix)
    lines = ['model_checkpoint_path: "%s"' % base]
    lines += ['all_model_checkpoint_paths: "%s"' % os.path.basename(p) for p in (keep or [prefix])]
    tmp = os.path.join(train_dir, "checkpoint.tmp")
    with open(tmp, "w") as f:
        f.write("\n".join(lines) + "\n")
    os.replace(tmp, os.path.join(train_dir, "checkpoint"))
