"""TFRecord datasets of point-wise data, column-sharded the way the reference writes them.

Reference: nif/data/tfr_dataset.py:8-163 (`TFRDataset`).  Each `.tfrecord` file holds ONE record, a
`tf.train.Example` whose features `input_0..`, `output_0..` (and `weight`) are FloatLists with one value per point
of that file ("I used an abnormal way to create tfrecord data", :139-141): the meta dataset yields one file at a
time and `gen_dataset_from_batch_file` re-batches its points.

TensorFlow is third-party there (`tf.io.TFRecordWriter`, `tf.train.Example`, `tf.data.TFRecordDataset`,
`tf.io.parse_single_example`) and absent here, so the two wire formats are restated from their published
definitions:

* TFRecord framing (tensorflow/core/lib/io/record_writer.h): per record
  `uint64 length | uint32 masked_crc32c(length) | bytes data | uint32 masked_crc32c(data)`, little endian,
  `masked(c) = ((c >> 15 | c << 17) + 0xa282ead8) mod 2^32`, CRC-32C (Castagnoli) from `nif_crc32c` in the C library.
* `tf.train.Example` (tensorflow/core/example/{example,feature}.proto), protobuf wire format:
  `Example{1: Features}`, `Features{1: repeated map entry {1: string key, 2: Feature}}`,
  `Feature{1: BytesList | 2: FloatList | 3: Int64List}`, `FloatList{1: repeated float, packed}`.

`tests/test_tfr_dataset.py` pins the CRC against the RFC 3720 vectors and the Example codec against the protobuf
runtime (message classes built from the same .proto definitions).
"""
from __future__ import annotations

import glob
import os
import struct
from typing import Dict, Iterator, List

import numpy as np

_MASK_DELTA = 0xA282EAD8


def crc32c(data: bytes, crc: int = 0) -> int:
    from .. import _lib
    return int(_lib.lib().nif_crc32c(data, len(data), crc))


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ---- TFRecord framing ---------------------------------------------------------------------------------
def write_tfrecord(path: str, records) -> None:
    with open(path, "wb") as f:
        for rec in records:
            head = struct.pack("<Q", len(rec))
            f.write(head)
            f.write(struct.pack("<I", masked_crc32c(head)))
            f.write(rec)
            f.write(struct.pack("<I", masked_crc32c(rec)))


def read_tfrecord(path: str, check_crc: bool = True) -> List[bytes]:
    out = []
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if not head:
                return out
            if len(head) != 8:
                raise ValueError(f"{path}: truncated record header")
            (n,) = struct.unpack("<Q", head)
            hcrc = f.read(4)
            data = f.read(n)
            dcrc = f.read(4)
            if len(hcrc) != 4 or len(data) != n or len(dcrc) != 4:
                raise ValueError(f"{path}: truncated record")
            if check_crc:
                if struct.unpack("<I", hcrc)[0] != masked_crc32c(head):
                    raise ValueError(f"{path}: corrupted record length")
                if struct.unpack("<I", dcrc)[0] != masked_crc32c(data):
                    raise ValueError(f"{path}: corrupted record data")
            out.append(data)


# ---- tf.train.Example with FloatList features ----------------------------------------------------------
def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:  # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features: Dict[str, np.ndarray]) -> bytes:
    """Serialise {name: 1-D float array} as tf.train.Example(features=Features(feature={name: FloatList})).
    Map entries are written in sorted key order (what the deterministic protobuf serialisation does)."""
    entries = b""
    for key in sorted(features):
        vals = np.ascontiguousarray(np.asarray(features[key], dtype="<f4").reshape(-1))
        float_list = _ld(1, vals.tobytes()) if vals.size else b""
        feature = _ld(2, float_list)
        entries += _ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feature))
    return _ld(1, entries)


def _fields(buf: bytes) -> Iterator:
    i, n = 0, len(buf)
    while i < n:
        tag = 0
        shift = 0
        while True:
            b = buf[i]
            i += 1
            tag |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v = 0
            shift = 0
            while True:
                b = buf[i]
                i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield field, wt, v
        elif wt == 1:
            yield field, wt, buf[i:i + 8]
            i += 8
        elif wt == 2:
            ln = 0
            shift = 0
            while True:
                b = buf[i]
                i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield field, wt, buf[i:i + ln]
            i += ln
        elif wt == 5:
            yield field, wt, buf[i:i + 4]
            i += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")


def decode_example(buf: bytes) -> Dict[str, np.ndarray]:
    """Parse the FloatList features of a serialised tf.train.Example (packed or unpacked floats)."""
    out: Dict[str, np.ndarray] = {}
    for f1, wt1, features in _fields(buf):
        if f1 != 1 or wt1 != 2:
            continue
        for f2, wt2, entry in _fields(features):
            if f2 != 1 or wt2 != 2:
                continue
            key, feature = None, b""
            for f3, wt3, v in _fields(entry):
                if f3 == 1 and wt3 == 2:
                    key = bytes(v).decode("utf-8")
                elif f3 == 2 and wt3 == 2:
                    feature = v
            if key is None:
                continue
            chunks = []
            for f4, wt4, lst in _fields(feature):
                if f4 != 2 or wt4 != 2:
                    continue  # bytes_list / int64_list: not used by TFRDataset
                for f5, wt5, v in _fields(lst):
                    if f5 == 1 and wt5 == 2:
                        chunks.append(np.frombuffer(bytes(v), dtype="<f4"))
                    elif f5 == 1 and wt5 == 5:
                        chunks.append(np.frombuffer(bytes(v), dtype="<f4"))
            out[key] = np.concatenate(chunks).astype(np.float32) if chunks else np.zeros(0, np.float32)
    return out


# ---- the reference's TFRDataset -------------------------------------------------------------------------
class TFRDataset(object):
    """Create and load TFRecord datasets (nif/data/tfr_dataset.py:8-163).

    Args:
        n_feature (int): number of input columns.
        n_target (int): number of target columns.
        area_weight (bool): the last column of the array is a per-point weight.
    """

    def __init__(self, n_feature, n_target, area_weight=False):
        self.n_feature = n_feature
        self.n_target = n_target
        self.area_weight = area_weight

    def _names(self) -> List[str]:
        names = ["input_" + str(j) for j in range(self.n_feature)] + ["output_" + str(j) for j in range(self.n_target)]
        return names + (["weight"] if self.area_weight else [])

    def create_from_npz(self, num_pts_per_file, npz_path, npz_key, tfr_path, prefix, seed=None):
        """Shuffle the rows, then write ceil(N / num_pts_per_file) files `{tfr_path}/{prefix}_{i}.tfrecord`, one
        Example each (tfr_dataset.py:22-88).  `seed` makes the shuffle reproducible (the reference uses the global
        numpy state)."""
        num_pts_per_file = int(num_pts_per_file)
        npz_data = np.array(np.load(npz_path)[npz_key])
        n_total, n_col = npz_data.shape
        assert n_col == self.n_feature + self.n_target + (1 if self.area_weight else 0)
        total_num_files = int(np.ceil(n_total / num_pts_per_file))
        (np.random.default_rng(seed) if seed is not None else np.random).shuffle(npz_data)
        os.makedirs(tfr_path, exist_ok=True)
        names = self._names()
        for i in range(total_num_files):
            rows = npz_data[i * num_pts_per_file:(i + 1) * num_pts_per_file]
            feats = {name: rows[:, j] for j, name in enumerate(names)}
            write_tfrecord(os.path.join(tfr_path, "{}_{}.tfrecord".format(prefix, i)), [encode_example(feats)])
        return total_num_files

    def get_tfr_meta_dataset(self, tfr_path, epoch, tfr_shuffle_buffer_size=1, seed=0):
        """Iterate over the files `epoch` times; every element is the list of that file's columns in schema order
        (input_0.., output_0.., weight), each shaped (1, n_points) like the reference's `.batch(1)`
        (tfr_dataset.py:126-163).  With tfr_shuffle_buffer_size > 1 the file order is shuffled every epoch."""
        filenames = sorted(glob.glob(os.path.join(tfr_path, "*.tfrecord")))
        self.num_pts_per_file = len(filenames)  # (sic) the reference stores the file count under this name
        names = self._names()
        rng = np.random.default_rng(seed)

        def gen():
            for _ in range(int(epoch)):
                order = list(filenames)
                if tfr_shuffle_buffer_size > 1:
                    rng.shuffle(order)
                for fn in order:
                    for rec in read_tfrecord(fn):
                        d = decode_example(rec)
                        missing = [k for k in names if k not in d]
                        if missing:
                            raise KeyError(f"{fn}: features {missing} are missing")
                        yield [d[k].reshape(1, -1) for k in names]
        return gen()

    def gen_dataset_from_batch_file(self, batch_file, batch_size):
        """One file's columns -> a shuffled, batched nif_b200.Dataset of (features, target[, weight])
        (tfr_dataset.py:90-124; targets are (n, n_target), without the reference's stray middle axis)."""
        from ..keras_like import Dataset
        cols = [np.asarray(c, np.float32).reshape(-1) for c in batch_file]
        features = np.stack(cols[: self.n_feature], 1)
        target = np.stack(cols[self.n_feature: self.n_feature + self.n_target], 1)
        arrays = [features, target]
        if self.area_weight:
            arrays.append(cols[-1].reshape(-1, 1))
        return Dataset(arrays).shuffle(features.shape[0]).batch(int(batch_size))
