"""Reader / writer for TensorFlow checkpoints (the TensorBundle format): what `model.save_weights("…/ckpt")` of the reference
writes and `load_weights` reads (tutorial/1_simple_1d_wave.ipynb:501, 1280; tutorial/2_multi_scale_NIF.ipynb:629), so that
weights trained with pswpswpsw/nif can be brought into nif_b200 and back.  TensorFlow itself is third-party there
(tensorflow==2.11.1, requirements.txt:3) and absent here; the format is restated from its published layout:

  <prefix>.data-00000-of-00001   the tensors' bytes, back to back
  <prefix>.index                 an SSTable (LevelDB table format) mapping
        ""                                  -> BundleHeaderProto  {num_shards, endianness, version}
        "<checkpoint key>"                  -> BundleEntryProto   {dtype, shape, shard_id, offset, size, masked crc32c}
        "_CHECKPOINTABLE_OBJECT_GRAPH"      -> a DT_STRING scalar holding a TrackableObjectGraph proto whose
                                               SerializedTensor records map each variable's full_name
                                               ("mlp_first_pnet/kernel", "HyperLinearForSIREN_w", …) to its checkpoint key
                                               ("layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE")
  table blocks: prefix-compressed entries + restart array, 1-byte compression type (0 raw, 1 snappy) + masked crc32c;
  footer: metaindex handle, index handle, padding to 40 bytes, magic 0xdb4775248b80fb57.

Host code: numpy + the library's crc32c (nif_crc32c).  No fixture of this format exists under /root/reference, so the reader is
pinned by round trips through the writer, hand-built snappy blocks and the published constants only.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

MAGIC = 0xDB4775248B80FB57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_, 14: None, 19: np.float16}
DTYPE_ID = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}
DT_STRING = 7


def _crc32c(data: bytes) -> int:
    from .tfr_dataset import crc32c
    return crc32c(data)


def _mask(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / protobuf wire format ----------------------------------------------------------------------------------
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    v = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field number, wire type, value)]: varint -> int, 64/32-bit -> bytes, length-delimited -> bytes."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((f, wt, v))
    return out


def _field(f: int, wt: int, payload) -> bytes:
    tag = _put_varint((f << 3) | wt)
    if wt == 0:
        return tag + _put_varint(int(payload))
    if wt == 2:
        return tag + _put_varint(len(payload)) + bytes(payload)
    if wt == 5:
        return tag + struct.pack("<I", int(payload))
    raise ValueError(wt)


# ---- snappy (block format) ----------------------------------------------------------------------------------------------
def snappy_decompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
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
            raise ValueError("corrupt snappy block")
        for _ in range(ln):  # overlapping copies repeat the pattern
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ---- SSTable ---------------------------------------------------------------------------------------------------------------
def _read_block(f: bytes, off: int, size: int, verify: bool = True) -> bytes:
    raw, ctype = f[off:off + size], f[off + size]
    if verify:
        want = struct.unpack("<I", f[off + size + 1:off + size + 5])[0]
        if _mask(_crc32c(f[off:off + size + 1])) != want:
            raise ValueError("checkpoint index: block checksum mismatch")
    if ctype == 0:
        return raw
    if ctype == 1:
        return snappy_decompress(raw)
    raise ValueError(f"checkpoint index: unknown block compression {ctype}")


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    nrestart = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * nrestart
    out, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(data: bytes) -> Dict[bytes, bytes]:
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != MAGIC:
        raise ValueError("not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _, p = _get_varint(footer, 0)      # metaindex handle
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    out: Dict[bytes, bytes] = {}
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        for k, v in _block_entries(_read_block(data, boff, bsize)):
            out[k] = v
    return out


def _build_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(items: Dict[bytes, bytes]) -> bytes:
    """One uncompressed data block + index block + empty metaindex (entries sorted by key)."""
    entries = sorted(items.items())
    out = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block)
        out.append(0)  # no compression
        out.extend(struct.pack("<I", _mask(_crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    data_handle = emit(_build_block(entries))
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block([(entries[-1][0] + b"\xff", data_handle)]))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    out.extend(footer)
    return bytes(out)


# ---- bundle protos --------------------------------------------------------------------------------------------------------
def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
    for f, wt, v in _parse_proto(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:  # Dim
                    size = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            size = v3
                    e["shape"].append(size)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
    return e


def _entry_proto(dtype_id: int, shape, offset: int, size: int, crc: int) -> bytes:
    shp = b"".join(_field(2, 2, _field(1, 0, int(d))) for d in shape)
    out = _field(1, 0, dtype_id) + _field(2, 2, shp)
    if offset:
        out += _field(4, 0, offset)
    return out + _field(5, 0, size) + _field(6, 5, crc)


def _parse_object_graph(buf: bytes) -> Dict[str, str]:
    """TrackableObjectGraph -> {variable full_name: checkpoint_key} (slot variables of optimisers included)."""
    names: Dict[str, str] = {}
    for f, _, node in _parse_proto(buf):
        if f != 1:
            continue
        for f2, _, attr in _parse_proto(node):
            if f2 != 2:
                continue
            rec = {1: b"", 2: b"", 3: b""}
            for f3, wt3, v3 in _parse_proto(attr):
                if f3 in rec and wt3 == 2:
                    rec[f3] = v3
            if rec[2] and rec[3]:
                names[rec[2].decode()] = rec[3].decode()
    return names


# ---- public API -------------------------------------------------------------------------------------------------------------
def read_checkpoint(prefix: str, verify: bool = True) -> Tuple[Dict[str, np.ndarray], Dict[str, str]]:
    """Returns ({checkpoint key: array}, {variable full_name: checkpoint key}) of `<prefix>.index` + its data shards."""
    with open(prefix + ".index", "rb") as fh:
        table = read_table(fh.read())
    header = table.pop(b"", None)
    num_shards = 1
    if header is not None:
        for f, _, v in _parse_proto(header):
            if f == 1:
                num_shards = v
            elif f == 2 and v != 0:
                raise ValueError("big-endian checkpoints are not supported")
    shards: Dict[int, bytes] = {}

    def shard(i: int) -> bytes:
        if i not in shards:
            with open(f"{prefix}.data-{i:05d}-of-{num_shards:05d}", "rb") as fh:
                shards[i] = fh.read()
        return shards[i]

    tensors: Dict[str, np.ndarray] = {}
    graph: Dict[str, str] = {}
    for key, val in table.items():
        e = _parse_entry(val)
        raw = shard(e["shard_id"])[e["offset"]:e["offset"] + e["size"]]
        name = key.decode()
        if e["dtype"] == DT_STRING:
            if name == OBJECT_GRAPH_KEY:  # scalar string: varint length, 4-byte checksum of the lengths, bytes
                ln, p = _get_varint(raw, 0)
                graph = _parse_object_graph(raw[p + 4:p + 4 + ln])
            continue
        if verify and e["crc32c"] is not None and _mask(_crc32c(raw)) != e["crc32c"]:
            raise ValueError(f"checkpoint tensor {name!r}: checksum mismatch")
        dt = DTYPES.get(e["dtype"])
        if dt is None:
            continue
        tensors[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    return tensors, graph


def load_variables(prefix: str) -> Dict[str, np.ndarray]:
    """{reference variable name: array}: the object graph's full_name where the checkpoint has one ("mlp_first_pnet/kernel",
    "siren_first_pnet_w", "HyperLinearForSIREN_b", …; optimiser slots are skipped), else the checkpoint key itself (name-based
    checkpoints)."""
    tensors, graph = read_checkpoint(prefix)
    out: Dict[str, np.ndarray] = {}
    if graph:
        for full_name, key in graph.items():
            if key in tensors and "/.OPTIMIZER_SLOT/" not in key and not key.startswith("optimizer/"):
                out[full_name.split(":")[0]] = tensors[key]
    else:
        for key, arr in tensors.items():
            out[key[:-len(SUFFIX)] if key.endswith(SUFFIX) else key] = arr
    return out


def write_checkpoint(prefix: str, variables: Dict[str, np.ndarray]) -> None:
    """Writes `<prefix>.index`, `<prefix>.data-00000-of-00001` and the `checkpoint` state file.  Variable i (in the given order)
    gets the key `variables/<i>/.ATTRIBUTES/VARIABLE_VALUE`, and the object graph records its name as full_name -- the layout
    of `tf.train.Checkpoint(variables=[...])`; readable by tf.train.load_checkpoint and by load_variables above.  (Keras
    `load_weights` additionally matches the model's own object graph, which only TensorFlow can produce.)"""
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    data = bytearray()
    items: Dict[bytes, bytes] = {}
    items[b""] = _field(1, 0, 1) + _field(3, 2, _field(1, 0, 1))  # num_shards = 1, little endian, version.producer = 1
    root_children, nodes = [], []
    for i, (name, arr) in enumerate(variables.items()):
        a = np.array(arr, order="C")  # (keeps 0-d arrays 0-d)
        if a.dtype not in DTYPE_ID:
            a = a.astype(np.float32)
        raw = a.tobytes()
        key = f"variables/{i}{SUFFIX}"
        items[key.encode()] = _entry_proto(DTYPE_ID[a.dtype], a.shape, len(data), len(raw), _mask(_crc32c(raw)))
        data += raw
        attr = _field(1, 2, b"VARIABLE_VALUE") + _field(2, 2, name.encode()) + _field(3, 2, key.encode())
        nodes.append(_field(2, 2, attr))
        root_children.append(_field(1, 2, _field(1, 0, i + 2) + _field(2, 2, str(i).encode())))
    # node 0: root -> "variables" (node 1) -> one node per variable
    graph = _field(1, 2, _field(1, 2, _field(1, 0, 1) + _field(2, 2, b"variables")))
    graph += _field(1, 2, b"".join(root_children))
    graph += b"".join(_field(1, 2, n) for n in nodes)
    lens = _put_varint(len(graph))
    sraw = lens + struct.pack("<I", _mask(_crc32c(lens))) + graph
    items[OBJECT_GRAPH_KEY.encode()] = _entry_proto(DT_STRING, (), len(data), len(sraw), _mask(_crc32c(sraw)))
    data += sraw
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))
    with open(prefix + ".index", "wb") as fh:
        fh.write(write_table(items))
    base = os.path.basename(prefix)
    with open(os.path.join(d or ".", "checkpoint"), "w") as fh:
        fh.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')
