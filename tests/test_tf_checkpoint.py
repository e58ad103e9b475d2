"""TensorFlow checkpoint (TensorBundle) reader / writer of nif_b200.data.tf_checkpoint: table format, protos, snappy blocks,
checksums, and name-matched loading into a model (CPU: no kernels are launched)."""
import os
import struct

import numpy as np
import pytest

import nif_b200
from nif_b200.data import tf_checkpoint as T


def test_varint_and_proto_roundtrip():
    for v in (0, 1, 127, 128, 300, 2**31, 2**40 + 5):
        assert T._get_varint(T._put_varint(v), 0) == (v, len(T._put_varint(v)))
    msg = T._field(1, 0, 7) + T._field(2, 2, b"abc") + T._field(6, 5, 0xDEADBEEF)
    assert T._parse_proto(msg) == [(1, 0, 7), (2, 2, b"abc"), (6, 5, struct.pack("<I", 0xDEADBEEF))]


def test_snappy_literals_and_copies():
    # "abcdabcdabcdX": literal "abcd", copy(offset 4, length 8) as a 2-byte-offset copy, literal "X"
    stream = T._put_varint(13) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((8 - 1) << 2) | 2, 4, 0]) + bytes([0]) + b"X"
    assert T.snappy_decompress(stream) == b"abcdabcdabcdX"
    # 1-byte-offset copy: length 4..11, offset < 2048:  tag = offset_hi << 5 | (len - 4) << 2 | 1
    stream = T._put_varint(10) + bytes([(5 - 1) << 2]) + b"hello" + bytes([((5 - 4) << 2) | 1, 5])
    assert T.snappy_decompress(stream) == b"hellohello"
    long_lit = bytes(range(256)) * 2  # literal of 512 bytes: length code 61 (2 extra bytes)
    stream = T._put_varint(512) + bytes([61 << 2]) + struct.pack("<H", 511) + long_lit
    assert T.snappy_decompress(stream) == long_lit
    with pytest.raises(ValueError):
        T.snappy_decompress(T._put_varint(3) + bytes([((4 - 4) << 2) | 1, 9]))  # copy before any output


def test_table_with_snappy_compressed_block(tmp_path):
    """A table whose data block is stored snappy-compressed (type byte 1), as TensorFlow writes large index blocks."""
    entries = [(b"", b"hdr"), (b"a/kernel", b"v1"), (b"a/kernel2", b"v2"), (b"b", b"v3")]
    block = T._build_block(entries, restart_interval=2)
    comp = T._put_varint(len(block)) + bytes([60 << 2, len(block) - 1]) + block  # one literal (1 extra length byte)
    assert len(block) <= 256
    out = bytearray(comp) + bytes([1])
    out += struct.pack("<I", T._mask(T._crc32c(bytes(out))))
    data_handle = T._put_varint(0) + T._put_varint(len(comp))
    meta_off = len(out)
    mb = T._build_block([])
    out += mb + b"\x00" + struct.pack("<I", T._mask(T._crc32c(mb + b"\x00")))
    idx_off = len(out)
    ib = T._build_block([(b"c", data_handle)])
    out += ib + b"\x00" + struct.pack("<I", T._mask(T._crc32c(ib + b"\x00")))
    footer = T._put_varint(meta_off) + T._put_varint(len(mb)) + T._put_varint(idx_off) + T._put_varint(len(ib))
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", T.MAGIC)
    assert T.read_table(bytes(out)) == dict(entries)


def test_checkpoint_roundtrip_and_checksums(tmp_path):
    rng = np.random.default_rng(0)
    variables = {"mlp_first_pnet/kernel": rng.normal(size=(1, 30)).astype(np.float32),
                 "mlp_first_pnet/bias": rng.normal(size=(30,)).astype(np.float32),
                 "HyperLinearForSIREN_w": rng.normal(size=(4, 1951)).astype(np.float32),
                 "scalar": np.float32(3.5).reshape(())}
    prefix = str(tmp_path / "ckpt-12" / "ckpt")
    T.write_checkpoint(prefix, variables)
    assert sorted(os.listdir(tmp_path / "ckpt-12")) == ["checkpoint", "ckpt.data-00000-of-00001", "ckpt.index"]
    tensors, graph = T.read_checkpoint(prefix)
    assert graph == {k: f"variables/{i}{T.SUFFIX}" for i, k in enumerate(variables)}
    got = T.load_variables(prefix)
    assert set(got) == set(variables)
    for k, v in variables.items():
        assert got[k].shape == v.shape and np.array_equal(got[k], v)
    # a flipped data byte is caught by the per-tensor crc32c, a flipped index byte by the block crc
    p = prefix + ".data-00000-of-00001"
    raw = bytearray(open(p, "rb").read())
    raw[10] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        T.read_checkpoint(prefix)
    raw[10] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[5] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        T.read_checkpoint(prefix)


def test_model_weights_through_a_tf_checkpoint(tmp_path):
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    a = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=1, device="cpu")
    b = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=2, device="cpu")
    prefix = str(tmp_path / "saved_weights" / "ckpt-4999" / "ckpt")  # the tutorials' path shape
    a.build().save_weights(prefix, save_format="tf")
    b.build().load_weights(prefix)
    for k, v in a.get_weights().items():
        assert np.array_equal(b.get_weights()[k], v), k
    # a checkpoint of a different architecture is refused with the missing names
    other = nif_b200.NIFMultiScale(dict(cfg_s, nlayers=3), dict(cfg_p, nlayers=3), seed=3, device="cpu")
    with pytest.raises(nif_b200._lib.NifError, match="lacks variables"):
        other.build().load_weights(prefix)


def test_keras_style_object_graph_names_are_resolved(tmp_path):
    """A checkpoint laid out the way Keras lays out a functional model -- keys layer_with_weights-N/<attr>/.ATTRIBUTES/
    VARIABLE_VALUE, names only in the object graph's full_name (with optimiser slots beside them)."""
    rng = np.random.default_rng(1)
    k0, b0 = rng.normal(size=(1, 4)).astype(np.float32), rng.normal(size=(4,)).astype(np.float32)
    slot = np.zeros((1, 4), np.float32)
    keys = {"layer_with_weights-0/kernel" + T.SUFFIX: k0, "layer_with_weights-0/bias" + T.SUFFIX: b0,
            "layer_with_weights-0/kernel/.OPTIMIZER_SLOT/optimizer/m" + T.SUFFIX: slot}
    names = {"mlp_first_pnet/kernel": "layer_with_weights-0/kernel" + T.SUFFIX,
             "mlp_first_pnet/bias": "layer_with_weights-0/bias" + T.SUFFIX,
             "Adam/mlp_first_pnet/kernel/m": "layer_with_weights-0/kernel/.OPTIMIZER_SLOT/optimizer/m" + T.SUFFIX}
    data, items = bytearray(), {b"": T._field(1, 0, 1)}
    for key, arr in keys.items():
        raw = arr.tobytes()
        items[key.encode()] = T._entry_proto(1, arr.shape, len(data), len(raw), T._mask(T._crc32c(raw)))
        data += raw
    node = b"".join(T._field(2, 2, T._field(1, 2, b"VARIABLE_VALUE") + T._field(2, 2, fn.encode()) + T._field(3, 2, ck.encode()))
                    for fn, ck in names.items())
    graph = T._field(1, 2, node)
    lens = T._put_varint(len(graph))
    sraw = lens + struct.pack("<I", T._mask(T._crc32c(lens))) + graph
    items[T.OBJECT_GRAPH_KEY.encode()] = T._entry_proto(T.DT_STRING, (), len(data), len(sraw), T._mask(T._crc32c(sraw)))
    data += sraw
    prefix = str(tmp_path / "ckpt")
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    open(prefix + ".index", "wb").write(T.write_table(items))
    got = T.load_variables(prefix)
    assert set(got) == {"mlp_first_pnet/kernel", "mlp_first_pnet/bias"}
    assert np.array_equal(got["mlp_first_pnet/kernel"], k0) and np.array_equal(got["mlp_first_pnet/bias"], b0)
