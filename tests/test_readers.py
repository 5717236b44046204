"""CPU suite: the TensorFlow-free YT8M frame-feature reader (readers.py:114-246 semantics).  The hand-written
protobuf wire codec is pinned against the protobuf library with tensorflow's example.proto / feature.proto
schema rebuilt at run time."""
import numpy as np
import pytest
import torch

from oracle import hlstm_oracle as O


def _tf_example_messages():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="tf_example_min.proto", package="tfmin", syntax="proto3")

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
        f = m.field.add(name=name, number=num, type=typ, label=label)
        if type_name:
            f.type_name = type_name
        if packed is not None:
            f.options.packed = packed
        if oneof is not None:
            f.oneof_index = oneof
        return f

    field(msg("BytesList"), "value", 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    field(msg("FloatList"), "value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    field(msg("Int64List"), "value", 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    feat = msg("Feature")
    feat.oneof_decl.add(name="kind")
    field(feat, "bytes_list", 1, F.TYPE_MESSAGE, type_name=".tfmin.BytesList", oneof=0)
    field(feat, "float_list", 2, F.TYPE_MESSAGE, type_name=".tfmin.FloatList", oneof=0)
    field(feat, "int64_list", 3, F.TYPE_MESSAGE, type_name=".tfmin.Int64List", oneof=0)

    def map_of(owner, fname, value_type):
        entry = owner.nested_type.add(name="".join(p.capitalize() for p in fname.split("_")) + "Entry")
        entry.options.map_entry = True
        field(entry, "key", 1, F.TYPE_STRING)
        field(entry, "value", 2, F.TYPE_MESSAGE, type_name=value_type)
        field(owner, fname, 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=f".tfmin.{owner.name}.{entry.name}")

    map_of(msg("Features"), "feature", ".tfmin.Feature")
    field(msg("FeatureList"), "feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tfmin.Feature")
    map_of(msg("FeatureLists"), "feature_list", ".tfmin.FeatureList")
    se = msg("SequenceExample")
    field(se, "context", 1, F.TYPE_MESSAGE, type_name=".tfmin.Features")
    field(se, "feature_lists", 2, F.TYPE_MESSAGE, type_name=".tfmin.FeatureLists")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tfmin.SequenceExample"))


def _video(rng, n, sizes):
    return {name: rng.integers(0, 256, size=(n, s), dtype=np.uint8) for name, s in sizes}


def test_wire_codec_against_protobuf_library():
    from efficientvideoclassification_youtube8m_b200 import readers
    SequenceExample = _tf_example_messages()
    rng = np.random.default_rng(0)
    feats = _video(rng, 7, [("rgb", 16), ("audio", 4)])
    ex = SequenceExample()
    ex.context.feature["id"].bytes_list.value.append(b"abcd")
    ex.context.feature["labels"].int64_list.value.extend([3, 17, 4000])
    for name, mat in feats.items():
        for row in mat:
            ex.feature_lists.feature_list[name].feature.add().bytes_list.value.append(row.tobytes())
    # library-serialised message parsed by the hand-written decoder
    ctx, lists = readers.parse_sequence_example(ex.SerializeToString())
    assert ctx["id"] == ("bytes", [b"abcd"]) and ctx["labels"] == ("int64", [3, 17, 4000])
    for name, mat in feats.items():
        got = np.stack([np.frombuffer(f[1][0], dtype=np.uint8) for f in lists[name]])
        assert np.array_equal(got, mat)
    # hand-written encoder parsed by the library
    ex2 = SequenceExample()
    ex2.ParseFromString(readers.make_sequence_example("abcd", [3, 17, 4000], feats))
    assert ex2 == ex


@pytest.mark.parametrize("native", [False, True])
def test_reader_batches_match_reference_semantics(tmp_path, native):
    from efficientvideoclassification_youtube8m_b200 import readers
    rng = np.random.default_rng(1)
    sizes = [("rgb", 32), ("audio", 8)]
    frames = [5, 300, 321, 1]                     # shorter, exact, longer than max_frames (truncated), single frame
    vids, recs = [], []
    for i, n in enumerate(frames):
        f = _video(rng, n, sizes)
        labels = sorted(rng.choice(50, size=3, replace=False).tolist())
        vids.append((f, labels))
        recs.append(readers.make_sequence_example(f"v{i}", labels, f))
    path = str(tmp_path / "train0.tfrecord")
    readers.write_tfrecord(path, recs)
    rd = readers.YT8MFrameFeatureReader(num_classes=50, feature_sizes=[32, 8], feature_names=["rgb", "audio"])
    out = list(rd.batches([path], batch_size=3, pin_memory=False, native=native))
    assert [len(b[0]) for b in out] == [3, 1]
    ids = sum((b[0] for b in out), [])
    x = torch.cat([b[1] for b in out]); y = torch.cat([b[2] for b in out]); nf = torch.cat([b[3] for b in out])
    assert ids == ["v0", "v1", "v2", "v3"]
    assert x.dtype == torch.uint8 and tuple(x.shape) == (4, 300, 40) and nf.dtype == torch.int32
    assert nf.tolist() == [5, 300, 300, 1]        # tf.minimum(num_frames, max_frames) (readers.py:168,237)
    for i, (f, labels) in enumerate(vids):
        k = min(frames[i], 300)
        want = np.concatenate([f["rgb"][:k], f["audio"][:k]], axis=1)     # tf.concat(feature_matrices, 1)
        assert np.array_equal(x[i, :k].numpy(), want)
        assert np.flatnonzero(y[i].numpy()).tolist() == labels
        # Dequantize + zero padding (utils.py:9-25, readers.py:173) = what the GPU pack kernel reproduces
        deq = np.where(np.arange(300)[:, None] < k, O.dequantize(x[i].numpy()), 0.0)
        assert deq[k:].sum() == 0 and abs(deq[:k].mean()) < 2.0
    with pytest.raises(AssertionError):
        readers.YT8MFrameFeatureReader(feature_sizes=[1024, 128], feature_names=["rgb"])


def _protobuf_record(SequenceExample, vid, labels, feats, packed_labels=True):
    ex = SequenceExample()
    ex.context.feature["id"].bytes_list.value.append(vid.encode())
    ex.context.feature["labels"].int64_list.value.extend(labels)
    for name, mat in feats.items():
        flist = ex.feature_lists.feature_list[name]          # present (and empty) also for a video without frames
        for row in mat:
            flist.feature.add().bytes_list.value.append(row.tobytes())
    return ex.SerializeToString()


def test_native_reader_equals_python_reader_on_library_serialised_shards(tmp_path):
    """libevc_reader (csrc/evc_reader.cpp) against the pure-Python decoder on records serialised by the
    protobuf LIBRARY (map entries, packed int64 labels, negative and out-of-range labels, an extra feature
    list the reader does not ask for), spread over three shards incl. an empty one, with TensorFlow-style
    CRCs verified; every thread count and batch size gives the same batches."""
    from efficientvideoclassification_youtube8m_b200 import readers
    SequenceExample = _tf_example_messages()
    rng = np.random.default_rng(3)
    sizes = [("rgb", 24), ("audio", 8), ("extra", 5)]
    shards = [[], [], []]
    for i in range(23):
        n = int(rng.integers(0, 330)) if i % 5 else (0 if i == 0 else 300)
        labels = rng.choice(60, size=int(rng.integers(0, 5)), replace=False).tolist() + ([-1, 50, 4096] if i % 4 == 0 else [])
        shards[0 if i < 9 else 2].append(_protobuf_record(SequenceExample, f"video-{i}", labels, _video(rng, n, sizes)))
    paths = []
    for k, recs in enumerate(shards):
        paths.append(str(tmp_path / f"train{k}.tfrecord"))
        readers.write_tfrecord(paths[-1], recs, with_crc=True)
    rd = readers.YT8MFrameFeatureReader(num_classes=50, feature_sizes=[24, 8], feature_names=["rgb", "audio"])
    want = list(rd.batches(paths, 4, pin_memory=False, native=False))
    assert sum(len(b[0]) for b in want) == 23
    for threads, prefetch in ((1, 0), (3, 0), (8, 2)):
        got = list((ids, x.clone(), y.clone(), n.clone()) for ids, x, y, n in
                   rd.batches(paths, 4, pin_memory=False, native=True, num_threads=threads, verify_crc=True,
                              prefetch=prefetch))
        assert len(got) == len(want)
        for (ia, xa, ya, na), (ib, xb, yb, nb) in zip(want, got):
            assert ia == ib and torch.equal(xa, xb) and torch.equal(ya, yb) and torch.equal(na, nb)
            assert yb.dtype == torch.bool and nb.dtype == torch.int32
    # drop_remainder and a batch size that does not divide the stream
    assert [len(b[0]) for b in rd.batches(paths, 5, pin_memory=False, drop_remainder=True)] == [5, 5, 5, 5]


def test_native_reader_errors(tmp_path):
    from efficientvideoclassification_youtube8m_b200 import readers
    rng = np.random.default_rng(4)
    rd = readers.YT8MFrameFeatureReader(num_classes=50, feature_sizes=[24, 8], feature_names=["rgb", "audio"])
    good = readers.make_sequence_example("ok", [1], _video(rng, 3, [("rgb", 24), ("audio", 8)]))
    # corrupted payload under CRC verification (TFRecordReader raises DataLossError)
    p = str(tmp_path / "a.tfrecord")
    readers.write_tfrecord(p, [good], with_crc=True)
    raw = bytearray(open(p, "rb").read())
    raw[40] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="CRC32C"):
        list(rd.batches([p], 2, pin_memory=False, verify_crc=True))
    # feature lists of different lengths (readers.py:218-219 asserts equal num_frames)
    bad = readers.make_sequence_example("bad", [1], {"rgb": _video(rng, 3, [("rgb", 24)])["rgb"],
                                                      "audio": _video(rng, 2, [("audio", 8)])["audio"]})
    p2 = str(tmp_path / "b.tfrecord")
    readers.write_tfrecord(p2, [good, bad])
    with pytest.raises(ValueError, match="frames"):
        list(rd.batches([p2], 2, pin_memory=False))
    with pytest.raises(ValueError):
        list(rd.batches([p2], 2, pin_memory=False, native=False))
    # a missing feature list, a truncated file, a missing file
    p3 = str(tmp_path / "c.tfrecord")
    readers.write_tfrecord(p3, [readers.make_sequence_example("x", [1], _video(rng, 3, [("rgb", 24)]))])
    with pytest.raises(ValueError, match="audio"):
        list(rd.batches([p3], 2, pin_memory=False))
    p4 = str(tmp_path / "d.tfrecord")
    open(p4, "wb").write(open(p2, "rb").read()[:-9])
    with pytest.raises(ValueError, match="truncated"):
        list(rd.batches([p4], 2, pin_memory=False))
    with pytest.raises(ValueError, match="cannot open"):
        list(rd.batches([str(tmp_path / "nope.tfrecord")], 2, pin_memory=False))


def test_masked_crc32c_known_answers():
    """CRC-32C check values (RFC 3720 B.4) through TensorFlow's mask: ((crc >> 15 | crc << 17) + 0xa282ead8)."""
    from efficientvideoclassification_youtube8m_b200 import readers
    lib = readers.reader_lib()

    def mask(c):
        return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF
    assert lib.evc_crc32c_masked(b"123456789", 9) == mask(0xE3069283)
    assert lib.evc_crc32c_masked(bytes(32), 32) == mask(0x8A9136AA)
    assert lib.evc_crc32c_masked(bytes([0xFF] * 32), 32) == mask(0x62A8AB43)
    assert lib.evc_crc32c_masked(bytes(range(32)), 32) == mask(0x46DD794E)


def test_get_input_data_batches_epochs_shuffle_and_rank_slices(tmp_path):
    """train.py:125-175: glob + IOError message, per-epoch file shuffle, num_epochs, smaller final batch; ranks
    read disjoint equal slices of the same permutation."""
    from efficientvideoclassification_youtube8m_b200 import readers
    rng = np.random.default_rng(9)
    for k in range(5):
        recs = [readers.make_sequence_example(f"s{k}v{i}", [k], _video(rng, 2, [("rgb", 8)])) for i in range(3)]
        readers.write_tfrecord(str(tmp_path / f"train{k}.tfrecord"), recs)
    rd = readers.YT8MFrameFeatureReader(num_classes=10, feature_sizes=[8], feature_names=["rgb"])
    pat = str(tmp_path / "train*.tfrecord")
    with pytest.raises(IOError, match="Unable to find training files. data_pattern="):
        next(readers.get_input_data_batches(rd, str(tmp_path / "nothing*")))
    ep = [b[0] for b in readers.get_input_data_batches(rd, pat, 4, num_epochs=2, pin_memory=False)]
    assert [len(b) for b in ep] == [4, 4, 4, 3, 4, 4, 4, 3]            # 15 videos per epoch
    e1, e2 = sum(ep[:4], []), sum(ep[4:], [])
    assert sorted(e1) == sorted(e2) and len(set(e1)) == 15 and e1 != e2      # reshuffled files, same videos
    for e in (e1, e2):                                                      # order inside a shard is kept
        for k in range(5):
            assert [v for v in e if v.startswith(f"s{k}")] == [f"s{k}v{i}" for i in range(3)]
    plain = sum((b[0] for b in readers.get_input_data_batches(rd, pat, 4, num_epochs=1, shuffle=False,
                                                               pin_memory=False)), [])
    assert plain == [f"s{k}v{i}" for k in range(5) for i in range(3)]
    # rank slices order[rank::world]: every shard is read by exactly one rank (the agreement on the epoch length
    # between ranks is covered with a real process group in tests/test_dp_gloo.py)
    seen = []
    for rank in range(2):
        ids = sum((b[0] for b in readers.get_input_data_batches(rd, pat, 4, num_epochs=1, rank=rank, world=2,
                                                                 pin_memory=False, agree=lambda have: (have, have))),
                  [])
        assert len(ids) == (9 if rank == 0 else 6)                          # 3 and 2 of the 5 shards
        seen.append(set(ids))
    assert not (seen[0] & seen[1]) and len(seen[0] | seen[1]) == 15
