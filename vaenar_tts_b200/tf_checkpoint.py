"""TensorFlow-free reader / writer of TF2 object-graph checkpoints (the "tensor bundle" format) for the VAENAR weights
-- SURVEY.md §8f rank 1: lets the checkpoints written by the reference (``tf.train.Checkpoint(step=, optimizer=, model=)``,
train.py:246-248, restored by inference.py:39-41,122-123) be loaded into ``vaenar_tts_b200.VAENAR.load_state_dict`` and
the trained weights + Adam state of this implementation be written back under the same keys.  No
``_CHECKPOINTABLE_OBJECT_GRAPH`` proto is written, so ``tf.train.Checkpoint.restore`` can NOT read the files this module
writes; ``tf.train.load_checkpoint(prefix).get_tensor(key)`` (name-based) can.

Format (public, stable since TF 1.x; restated from tensorflow/core/util/tensor_bundle and the LevelDB table format):

  <prefix>.index                  an SSTable: sorted key -> value, blocks of prefix-compressed entries
                                  [shared varint32][non_shared varint32][value_len varint32][key suffix][value],
                                  restart array, 1-byte compression type + 4-byte masked CRC32C per block,
                                  48-byte footer (metaindex handle, index handle, magic 0xdb4775248b80fb57).
                                  key ""  -> BundleHeaderProto {num_shards=1, endianness=2, version=3}
                                  key k   -> BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}
  <prefix>.data-00000-of-00001    raw little-endian tensor bytes at [offset, offset + size)

Object-graph keys: ``model/<attribute path>/.ATTRIBUTES/VARIABLE_VALUE``; list elements and the (actnorm, linear,
affine_coupling) tuples of ``prior.glow`` (modules/prior.py:84-99) appear as integer path components.

PARITY UNPINNED: TensorFlow is not installable in the build image and the reference ships no checkpoint, so this module
is pinned only by the format's known-answer vectors (CRC32C, varints, the table magic) and by write -> read round trips
(tests/test_tf_checkpoint_cpu.py).  Snappy-compressed index blocks (never written by TF's BundleWriter) are rejected.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_GLOW = ("actnorm", "linear", "affine_coupling")
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 10: np.dtype("bool")}
_DTYPE_ENUM = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9}


# ------------------------------------------------------------------------------------------------ CRC32C (Castagnoli)
def _make_table():
    poly = 0x82F63B78
    tbl = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tbl[i] = c
    return tbl


_TBL = _make_table()
_TBL_LIST = [int(x) for x in _TBL]


def _native_crc():
    try:
        from . import _lib
        return _lib.load().vaenar_crc32c
    except Exception:
        return None


def crc32c(data: bytes, crc: int = 0) -> int:
    fn = _native_crc() if len(data) > 4096 else None      # the C-ABI library carries a table-driven host implementation
    if fn is not None:
        import ctypes
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
        return int(fn(ctypes.cast(buf, ctypes.c_void_p), len(data), crc))
    c = crc ^ 0xFFFFFFFF
    t = _TBL_LIST
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    """The 'masked' CRC stored in tables and TFRecords: rotate right by 15 and add a constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """Generic protobuf scan: {field number: [values]} with varints as int, length-delimited as bytes, fixed32/64 as int."""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _entry_proto(dtype_enum, shape, offset, size, crc):
    dims = b"".join(b"\x12" + _varint(len(d)) + d for d in (b"\x08" + _varint(int(s)) for s in shape))
    msg = b"\x08" + _varint(dtype_enum) + b"\x12" + _varint(len(dims)) + dims
    if offset:
        msg += b"\x20" + _varint(offset)
    msg += b"\x28" + _varint(size) + b"\x35" + struct.pack("<I", crc)
    return msg


# ------------------------------------------------------------------------------------------------ SSTable
def _read_block(buf, offset, size):
    block = buf[offset:offset + size]
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed index block (type %d): TF's BundleWriter writes uncompressed tables" % ctype)
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if masked_crc32c(bytes(block) + bytes([ctype])) != stored:
        raise ValueError("index block checksum mismatch")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    entries, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        entries.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return entries


def read_table(path):
    """All (key, value) pairs of an SSTable file, in key order."""
    buf = memoryview(open(path, "rb").read())
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TensorFlow/LevelDB table (bad magic)")
    footer = buf[len(buf) - 48:]
    _, pos = _read_varint(footer, 0)          # metaindex handle
    _, pos = _read_varint(footer, pos)
    idx_off, pos = _read_varint(footer, pos)
    idx_size, pos = _read_varint(footer, pos)
    out = []
    for _, handle in _read_block(buf, idx_off, idx_size):
        off, p = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p)
        out.extend(_read_block(buf, off, size))
    return out


def _build_block(entries, restart_interval=16):
    body, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def write_table(path, items, block_entries=64):
    """Write sorted (key, value) pairs as an uncompressed SSTable."""
    items = sorted(items)
    out = bytearray()
    index = []

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return off, len(block)
    for i in range(0, len(items), block_entries):
        chunk = items[i:i + block_entries]
        off, size = emit(_build_block(chunk))
        index.append((chunk[-1][0], _varint(off) + _varint(size)))      # separator key >= every key of the block
    if not index:
        off, size = emit(_build_block([]))
        index.append((b"", _varint(off) + _varint(size)))
    meta_off, meta_size = emit(_build_block([]))
    idx_off, idx_size = emit(_build_block(index, restart_interval=1))
    footer = _varint(meta_off) + _varint(meta_size) + _varint(idx_off) + _varint(idx_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(out)


# ------------------------------------------------------------------------------------------------ tensor bundle
def read_bundle(prefix, verify_crc_below=1 << 40):
    """{key: numpy array} of every numeric tensor of the checkpoint ``prefix`` (.index + .data-*).  The CRC32C of tensors
    smaller than ``verify_crc_below`` bytes is verified."""
    entries = read_table(prefix + ".index")
    header = None
    out = {}
    shards = {}
    for key, val in entries:
        if key == b"":
            header = _parse_proto(val)
            continue
        e = _parse_proto(val)
        dt = e.get(1, [0])[0]
        if dt not in _DTYPES or 7 in e:          # strings (the object graph proto) and sliced tensors: not weights
            continue
        shape = []
        if 2 in e:
            for dim in _parse_proto(e[2][0]).get(2, []):
                shape.append(_parse_proto(dim).get(1, [0])[0])
        shard, offset, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        n_shards = header.get(1, [1])[0] if header else 1
        if shard not in shards:
            shards[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, n_shards), dtype=np.uint8, mode="r")
        raw = bytes(shards[shard][offset:offset + size])
        if size < verify_crc_below and 6 in e and masked_crc32c(raw) != e[6][0]:
            raise ValueError(f"{key.decode()}: tensor checksum mismatch")
        out[key.decode()] = np.frombuffer(raw, dtype=_DTYPES[dt]).reshape(shape).copy()
    if header is None:
        raise ValueError(f"{prefix}.index: no bundle header")
    if header.get(2, [0])[0] == 1:
        raise ValueError("big-endian bundles are not supported")
    return out


def write_bundle(prefix, tensors):
    """Write {key: numpy array} as a single-shard tensor bundle."""
    data = bytearray()
    items = [(b"", b"\x08\x01\x1a\x02\x08\x01")]            # num_shards = 1, endianness LITTLE (0, default), version {producer: 1}
    for key in sorted(tensors):
        a = np.asarray(tensors[key], order="C")
        if a.dtype not in _DTYPE_ENUM:
            raise ValueError(f"{key}: unsupported dtype {a.dtype}")
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        items.append((key.encode(), _entry_proto(_DTYPE_ENUM[a.dtype], a.shape, len(data), len(raw), masked_crc32c(raw))))
        data.extend(raw)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data)
    write_table(prefix + ".index", items)


# ------------------------------------------------------------------------------------------------ name mapping
def tf_key_to_name(key: str):
    """``model/prior/glow/3/2/net/attentions/1/ffn/dense2/kernel/.ATTRIBUTES/VARIABLE_VALUE`` ->
    ``prior.glow.3.affine_coupling.net.attentions.1.ffn.dense2.kernel``; None for keys that are not model weights
    (optimizer slots, the step counter, the object graph)."""
    if not key.startswith("model/") or not key.endswith(_SUFFIX) or "/.OPTIMIZER_SLOT/" in key:
        return None
    parts = key[len("model/"):-len(_SUFFIX)].split("/")
    out = []
    for i, p in enumerate(parts):
        if len(out) >= 3 and out[-3] == "prior" and out[-2] == "glow" and out[-1].isdigit() and p.isdigit():
            p = _GLOW[int(p)]                      # the (actnorm, linear, coupling) tuple of modules/prior.py:84-99
        out.append(p)
    return ".".join(out)


def name_to_tf_key(name: str):
    parts = name.split(".")
    out = []
    for i, p in enumerate(parts):
        if len(out) >= 3 and out[-3] == "prior" and out[-2] == "glow" and out[-1].isdigit() and p in _GLOW:
            p = str(_GLOW.index(p))
        out.append(p)
    return "model/" + "/".join(out) + _SUFFIX


def load_tf_checkpoint(prefix):
    """State dict {attribute.path: numpy array} of the model weights stored in the TF checkpoint ``prefix``
    (e.g. ``.../ckpt-20``), ready for ``VAENAR.load_state_dict(..., strict=False)``."""
    sd = {}
    for key, arr in read_bundle(prefix).items():
        name = tf_key_to_name(key)
        if name is not None:
            sd[name] = arr
    return sd


def _slot_key(name: str, slot: str):
    """Keras optimizer slot of a model variable in the object graph: ``model/<path>/.OPTIMIZER_SLOT/optimizer/<m|v>/...``"""
    return name_to_tf_key(name)[:-len(_SUFFIX)] + "/.OPTIMIZER_SLOT/optimizer/" + slot + _SUFFIX


def save_tf_checkpoint(prefix, state_dict, step=0, adam_m=None, adam_v=None, opt_step=None):
    """Write the model weights under the reference's object-graph keys, the ``step`` counter of train.py:246 and -- when
    given -- the Adam moments (``{name: tensor}``, slots ``m`` / ``v``) and the optimizer's iteration count
    (``optimizer/iter``), i.e. everything ``tf.train.Checkpoint(step=, optimizer=, model=)`` (train.py:246) persists, so a
    resumed run continues with the right bias correction instead of restarting Adam at t = 1.  Scalars (``pos_weight``,
    ``step``, ``iter``) are written with shape [] like TF does.  No object-graph proto is written (see module docstring)."""
    tensors = {}
    for k, v in state_dict.items():
        a = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)
        if k.endswith("pos_weight"):
            a = a.reshape(())                       # tf.Variable(1.0): shape [] (encoder.py:64, posterior.py:95, transform.py:35)
        tensors[name_to_tf_key(k)] = a
    for slot, src in (("m", adam_m), ("v", adam_v)):
        if src is None:
            continue
        for k, v in src.items():
            a = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)
            if k.endswith("pos_weight"):
                a = a.reshape(())
            tensors[_slot_key(k, slot)] = a
    if opt_step is not None:
        tensors["optimizer/iter" + _SUFFIX] = np.asarray(int(opt_step), dtype=np.int64)
    tensors["step" + _SUFFIX] = np.asarray(step, dtype=np.int64)
    write_bundle(prefix, tensors)


def load_tf_optimizer_state(prefix):
    """(adam_m, adam_v, iter) stored in the checkpoint ``prefix`` ({name: array} per slot; iter None when absent)."""
    m, v, it = {}, {}, None
    for key, a in read_bundle(prefix).items():
        if key == "optimizer/iter" + _SUFFIX:
            it = int(a)
        elif "/.OPTIMIZER_SLOT/optimizer/" in key and key.endswith(_SUFFIX):
            base, slot = key[:-len(_SUFFIX)].split("/.OPTIMIZER_SLOT/optimizer/")
            name = tf_key_to_name(base + _SUFFIX)
            if name is not None and slot in ("m", "v"):
                (m if slot == "m" else v)[name] = a
    return m, v, it
