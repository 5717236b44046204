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


def test_reader_batches_match_reference_semantics(tmp_path):
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
    out = list(rd.batches([path], batch_size=3, pin_memory=False))
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
