"""TFRecord / tf.train.Example codec and the TFRDataset mirror (reference: nif/data/tfr_dataset.py).

TensorFlow is absent, so the codec is pinned against (a) the RFC 3720 B.4 CRC-32C vectors and the masked-CRC definition of
the TFRecord format, (b) the protobuf runtime with message classes built from the tf.train.Example definitions
(tensorflow/core/example/{example,feature}.proto)."""
import os
import struct

import numpy as np
import pytest

from nif_b200.data import TFRDataset
from nif_b200.data import tfr_dataset as T


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA            # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43          # 32 bytes of ones
    assert T.crc32c(bytes(range(32))) == 0x46DD794E      # incrementing
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C  # decrementing
    data = os.urandom(1000)
    assert T.crc32c(data[400:], T.crc32c(data[:400])) == T.crc32c(data)  # continuation
    assert T.crc32c(b"") == 0


def _example_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="nif_b200_test_example.proto", package="tfx", syntax="proto3")
    L = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    m = msg("BytesList")
    m.field.add(name="value", number=1, type=L.TYPE_BYTES, label=L.LABEL_REPEATED)
    m = msg("FloatList")
    m.field.add(name="value", number=1, type=L.TYPE_FLOAT, label=L.LABEL_REPEATED)
    m = msg("Int64List")
    m.field.add(name="value", number=1, type=L.TYPE_INT64, label=L.LABEL_REPEATED)
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    m.field.add(name="bytes_list", number=1, type=L.TYPE_MESSAGE, type_name=".tfx.BytesList", label=L.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="float_list", number=2, type=L.TYPE_MESSAGE, type_name=".tfx.FloatList", label=L.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="int64_list", number=3, type=L.TYPE_MESSAGE, type_name=".tfx.Int64List", label=L.LABEL_OPTIONAL, oneof_index=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry")
    e.options.map_entry = True
    e.field.add(name="key", number=1, type=L.TYPE_STRING, label=L.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=L.TYPE_MESSAGE, type_name=".tfx.Feature", label=L.LABEL_OPTIONAL)
    m.field.add(name="feature", number=1, type=L.TYPE_MESSAGE, type_name=".tfx.Features.FeatureEntry", label=L.LABEL_REPEATED)
    m = msg("Example")
    m.field.add(name="features", number=1, type=L.TYPE_MESSAGE, type_name=".tfx.Features", label=L.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:
        return message_factory.MessageFactory(pool).GetPrototype(pool.FindMessageTypeByName("tfx.Example"))
    return get(pool.FindMessageTypeByName("tfx.Example"))


def test_example_codec_matches_protobuf_runtime():
    Example = _example_classes()
    rng = np.random.default_rng(0)
    feats = {"input_0": rng.normal(size=300).astype(np.float32), "input_1": rng.normal(size=300).astype(np.float32),
             "output_0": rng.normal(size=300).astype(np.float32), "weight": rng.uniform(size=300).astype(np.float32)}
    # ours -> protobuf
    ex = Example()
    ex.ParseFromString(T.encode_example(feats))
    assert sorted(ex.features.feature.keys()) == sorted(feats)
    for k, v in feats.items():
        assert np.array_equal(np.asarray(ex.features.feature[k].float_list.value, np.float32), v)
    # protobuf -> ours (what tf.train.Example(...).SerializeToString() produces)
    ex2 = Example()
    for k, v in feats.items():
        ex2.features.feature[k].float_list.value.extend(v.tolist())
    got = T.decode_example(ex2.SerializeToString())
    assert sorted(got) == sorted(feats)
    for k, v in feats.items():
        assert np.array_equal(got[k], v)
    # and the deterministic serialisation is byte-identical to ours
    assert ex2.SerializeToString(deterministic=True) == T.encode_example(feats)


def test_tfrecord_framing_and_corruption(tmp_path):
    recs = [b"", b"abc", os.urandom(5000)]
    p = str(tmp_path / "a.tfrecord")
    T.write_tfrecord(p, recs)
    assert T.read_tfrecord(p) == recs
    raw = open(p, "rb").read()
    # layout of the first (empty) record: length 0, masked crc of the length, no data, masked crc of b""
    assert struct.unpack("<Q", raw[:8])[0] == 0
    assert struct.unpack("<I", raw[8:12])[0] == T.masked_crc32c(raw[:8])
    assert struct.unpack("<I", raw[12:16])[0] == ((((0 >> 15) | (0 << 17)) + 0xA282EAD8) & 0xFFFFFFFF)
    bad = bytearray(raw)
    bad[-10] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        T.read_tfrecord(p)
    assert len(T.read_tfrecord(p, check_crc=False)) == 3
    open(p, "wb").write(raw[:-3])
    with pytest.raises(ValueError):
        T.read_tfrecord(p)


@pytest.mark.parametrize("area_weight", [False, True])
def test_tfrdataset_round_trip(tmp_path, area_weight):
    rng = np.random.default_rng(1)
    n, nf, nt = 1000, 3, 2
    data = rng.normal(size=(n, nf + nt + (1 if area_weight else 0))).astype(np.float32)
    npz = str(tmp_path / "d.npz")
    np.savez(npz, data=data)
    ds = TFRDataset(nf, nt, area_weight)
    nfiles = ds.create_from_npz(300, npz, "data", str(tmp_path / "tfr"), "case", seed=3)
    assert nfiles == 4 and sorted(os.listdir(tmp_path / "tfr")) == [f"case_{i}.tfrecord" for i in range(4)]
    rows = []
    sizes = []
    for batch_file in ds.get_tfr_meta_dataset(str(tmp_path / "tfr"), epoch=1):
        assert len(batch_file) == nf + nt + (1 if area_weight else 0)
        assert all(c.shape == (1, batch_file[0].shape[1]) for c in batch_file)
        sizes.append(batch_file[0].shape[1])
        sub = ds.gen_dataset_from_batch_file(batch_file, 128)
        got = 0
        for gb, parts in sub.batches(0):
            assert parts[0].shape[1] == nf and parts[1].shape[1] == nt
            if area_weight:
                assert parts[2].shape[1] == 1
            rows.append(np.hstack([np.asarray(p) for p in parts]))
            got += gb
        assert got == batch_file[0].shape[1]
    assert sorted(sizes) == [100, 300, 300, 300]
    allrows = np.vstack(rows)
    # every point comes back exactly once (the files and the sub-batches are shuffled)
    key = lambda a: a[np.lexsort(a.T[::-1])]
    assert np.array_equal(key(allrows), key(data))
    assert ds.num_pts_per_file == 4
    assert sum(1 for _ in ds.get_tfr_meta_dataset(str(tmp_path / "tfr"), epoch=3, tfr_shuffle_buffer_size=4)) == 12
