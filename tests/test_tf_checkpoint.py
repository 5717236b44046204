"""TF-checkpoint (tensor bundle V2) reader/writer without TensorFlow (CPU).  No TensorFlow exists here to produce
a reference file, so the reader is pinned against hand-assembled tables / protos written byte by byte from the
published format (prefix compression, restart points, snappy blocks, two data shards), against published CRC32C
vectors, and against the writer."""
import os
import struct

import numpy as np
import pytest

from efficientvideoclassification_youtube8m_b200 import tf_checkpoint as C


def test_crc32c_vectors_and_mask():
    # RFC 3720 B.4 test vectors
    assert C.crc32c(b"123456789") == 0xE3069283
    assert C.crc32c(bytes(32)) == 0x8A9136AA
    assert C.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    big = bytes(range(256)) * 64                       # > 4096 bytes: the native path, same polynomial
    c = 0xFFFFFFFF
    for b in big:
        c = int(C._CRC_TABLE[(c ^ b) & 0xFF]) ^ (c >> 8)
    assert C.crc32c(big) == c ^ 0xFFFFFFFF
    for v in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert C._unmask(C._mask(v)) == v
    assert C._mask(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _block(entries, restarts):
    body = b""
    for shared, key_delta, value in entries:
        body += bytes([shared, len(key_delta), len(value)]) + key_delta + value
    return body + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def _with_trailer(block, ctype=0):
    return block + bytes([ctype]) + struct.pack("<I", C._mask(C.crc32c(block + bytes([ctype]))))


def test_hand_assembled_table_prefix_compression_restarts_and_snappy():
    # data block 0: keys "model/a", "model/ab" (shares 7), restart, "model/b"
    e0 = [(0, b"model/a", b"v1"), (7, b"b", b"v2"), (0, b"model/b", b"v3")]
    b0 = _block(e0, [0, len(bytes([0, 7, 2]) + b"model/a" + b"v1" + bytes([7, 1, 2]) + b"b" + b"v2")])
    # data block 1, snappy: one literal element holding the whole block
    raw1 = _block([(0, b"model/c", b"x" * 40), (6, b"d", b"")], [0])
    snap = bytes([len(raw1)]) + bytes([60 << 2, len(raw1) - 1]) + raw1          # varint length, literal tag 60 = 1-byte length
    file = bytearray()
    h0 = (len(file), len(b0)); file += _with_trailer(b0)
    h1 = (len(file), len(snap)); file += _with_trailer(snap, 1)
    meta = _block([], [0]); hm = (len(file), len(meta)); file += _with_trailer(meta)
    idx = _block([(0, b"model/b", bytes(h0)), (0, b"model/z", bytes(h1))], [0, 3 + 7 + 2])
    hi = (len(file), len(idx)); file += _with_trailer(idx)
    footer = (bytes(hm) + C._put_varint(hi[0]) + C._put_varint(hi[1])).ljust(40, b"\0") + struct.pack("<Q", C.TABLE_MAGIC)
    pairs = C.read_table(bytes(file) + footer)
    assert pairs == [(b"model/a", b"v1"), (b"model/ab", b"v2"), (b"model/b", b"v3"), (b"model/c", b"x" * 40),
                     (b"model/d", b"")]
    bad = bytearray(bytes(file) + footer)
    bad[3] ^= 1
    with pytest.raises(ValueError, match="checksum"):
        C.read_table(bytes(bad))
    with pytest.raises(ValueError, match="magic"):
        C.read_table(bytes(file) + footer[:-1] + b"\0")


def test_snappy_copy_elements():
    # "abcdabcdabcdabcd!": literal "abcd", 1-byte-offset copies (len 8 [overlapping its own output], len 4), literal "!"
    s = (bytes([17]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([((4 - 4) << 2) | 1, 4]) +
         bytes([0]) + b"!")
    assert C.snappy_decompress(s) == b"abcdabcdabcdabcd!"
    s2 = bytes([8]) + bytes([3 << 2]) + b"wxyz" + bytes([((4 - 1) << 2) | 2, 4, 0])       # 2-byte-offset copy
    assert C.snappy_decompress(s2) == b"wxyzwxyz"
    with pytest.raises(ValueError):
        C.snappy_decompress(bytes([5]) + bytes([((4 - 1) << 2) | 2, 9, 0]))


def test_bundle_hand_assembled_two_shards(tmp_path):
    """BundleHeaderProto / BundleEntryProto written byte by byte, tensors spread over two data shards."""
    prefix = str(tmp_path / "model.ckpt-7")
    a = np.arange(6, dtype="<f4").reshape(2, 3)
    g = np.array(7, dtype="<i8")
    with open(prefix + ".data-00000-of-00002", "wb") as f:
        f.write(b"\0" * 8 + a.tobytes())               # tensor at offset 8
    with open(prefix + ".data-00001-of-00002", "wb") as f:
        f.write(g.tobytes())
    ent_a = (b"\x08\x01" + b"\x12\x08" + b"\x12\x02\x08\x02" + b"\x12\x02\x08\x03" + b"\x20\x08" + b"\x28\x18" +
             b"\x35" + struct.pack("<I", C._mask(C.crc32c(a.tobytes()))))
    ent_g = (b"\x08\x09" + b"\x12\x00" + b"\x18\x01" + b"\x28\x08" + b"\x35" + struct.pack("<I", C._mask(C.crc32c(g.tobytes()))))
    header = b"\x08\x02\x1a\x02\x08\x01"              # num_shards 2, little endian (default), version.producer 1
    with open(prefix + ".index", "wb") as f:
        f.write(C.write_table([(b"", header), (b"global_step", ent_g), (b"model/w", ent_a)]))
    assert C.list_variables(prefix) == {"global_step": (np.dtype("<i8"), ()), "model/w": (np.dtype("<f4"), (2, 3))}
    got = C.load_variables(prefix)
    assert np.array_equal(got["model/w"], a) and got["global_step"] == 7 and got["global_step"].shape == ()
    assert list(C.load_variables(prefix, ["model/w"])) == ["model/w"]
    with pytest.raises(KeyError):
        C.load_variables(prefix, ["model/missing"])
    with open(prefix + ".data-00000-of-00002", "r+b") as f:      # flip a tensor byte: per-tensor crc catches it
        f.seek(9)
        f.write(b"\x55")
    with pytest.raises(ValueError, match="checksum"):
        C.load_variables(prefix)


def test_writer_reader_roundtrip_many_blocks(tmp_path):
    """The 22 variables of a teacher+student checkpoint (+ Adam slots, global_step): > 1 index block, names with
    long shared prefixes; latest_checkpoint / CheckpointState file."""
    rng = np.random.default_rng(0)
    names = []
    for scope in ("model", "model_student"):
        for level in ("RNN_L1", "RNN_L2"):
            for cell in (0, 1):
                base = f"{scope}/{level}/rnn/multi_rnn_cell/cell_{cell}/basic_lstm_cell"
                names += [base + "/kernel", base + "/bias"]
        names += [f"{scope}/classifier/gates/weights", f"{scope}/classifier/experts/weights",
                  f"{scope}/classifier/experts/biases"]
    var = {}
    for n in names:
        shape = (rng.integers(1, 9),) if n.endswith(("bias", "biases")) else (rng.integers(1, 9), rng.integers(1, 9))
        var[n] = rng.standard_normal(shape).astype(np.float32)
        var[n + "/Adam"] = np.zeros(shape, np.float32)
        var[n + "/Adam_1"] = np.ones(shape, np.float32)
    var["global_step"] = np.array(36704, dtype=np.int64)
    var["beta1_power"] = np.array(0.9, dtype=np.float32)
    prefix = str(tmp_path / "model.ckpt-36704")
    C.save_variables(prefix, var)
    pairs = C.read_table(open(prefix + ".index", "rb").read())
    small = C.write_table(pairs, block_size=256, restart_interval=3)           # many data blocks, many restarts
    assert C.read_table(small) == pairs and len(small) > os.path.getsize(prefix + ".index")
    got = C.load_variables(prefix)
    assert set(got) == set(var)
    for n in var:
        assert got[n].dtype == var[n].dtype and np.array_equal(got[n], var[n]), n
    assert C.latest_checkpoint(str(tmp_path)) is None
    C.update_checkpoint_state(str(tmp_path), prefix)
    assert C.latest_checkpoint(str(tmp_path)) == prefix
    assert open(tmp_path / "checkpoint").read().startswith('model_checkpoint_path: "model.ckpt-36704"\n')
    # keys are sorted bytewise and prefix-compressed: the raw index is much smaller than the sum of the names
    raw = open(prefix + ".index", "rb").read()
    assert raw.count(b"multi_rnn_cell") < len(names)
