"""TensorFlow-free input pipeline for the records the reference writes:
``datasets/tf_record_utils.py`` -- ``TFRecordWriter.serialize_example`` (:35-54), ``parse_example`` (:108-124) and
``create_dataset`` (:126-142: TFRecordDataset -> map(parse) -> padded_batch -> shuffle of batches -> prefetch).

  * TFRecord framing: [uint64 length][uint32 masked_crc32c(length)][payload][uint32 masked_crc32c(payload)]
  * payload = ``tf.train.Example``: features['fid'] bytes, ['text'] / ['mel'] = ``tf.io.serialize_tensor`` (a TensorProto
    with dtype / tensor_shape / tensor_content: int64 ids, float64 [mel_len, num_mels]), ['text_len'] / ['mel_len'] int64
  * batches: zero-padded to the longest item (``padded_shapes=([], [None], [None, num_mels], [], [])``), texts / lengths
    cast to int32, mels to float32 (tf_record_utils.py:124), delivered as PINNED torch tensors ready for
    ``VAENAR.train_step`` (the host->device copy is then asynchronous).

Reading order: ``TFRecordDataset(files, num_parallel_reads=n)`` is a deterministic interleave with cycle length n and
block length 1; ``iter_records`` reproduces that order.  ``shuffle`` permutes BATCHES through a buffer like
``Dataset.shuffle`` does, but with numpy's generator: the same distribution, not TensorFlow's permutation.

PARITY UNPINNED against files written by TensorFlow (none can be produced in this image): pinned by the CRC32C test
vectors, a hand-encoded Example, and write -> read round trips (tests/test_data_cpu.py).
"""
import struct

import numpy as np
import torch

from .tf_checkpoint import _parse_proto, _varint, masked_crc32c

_DT_FLOAT, _DT_DOUBLE, _DT_INT32, _DT_INT64 = 1, 2, 3, 9
_NP = {_DT_FLOAT: "<f4", _DT_DOUBLE: "<f8", _DT_INT32: "<i4", _DT_INT64: "<i8"}
_ENUM = {np.dtype("float32"): _DT_FLOAT, np.dtype("float64"): _DT_DOUBLE, np.dtype("int32"): _DT_INT32,
         np.dtype("int64"): _DT_INT64}


# ------------------------------------------------------------------------------------------------ TFRecord framing
def read_tfrecord(path, verify_crc=True):
    """Yield the payload of every record of one .tfrecords file."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError(f"{path}: truncated record header")
            n, = struct.unpack("<Q", head[:8])
            if verify_crc and masked_crc32c(head[:8]) != struct.unpack("<I", head[8:])[0]:
                raise ValueError(f"{path}: corrupted record length")
            data = f.read(n)
            foot = f.read(4)
            if len(data) < n or len(foot) < 4:
                raise ValueError(f"{path}: truncated record")
            if verify_crc and masked_crc32c(data) != struct.unpack("<I", foot)[0]:
                raise ValueError(f"{path}: corrupted record payload")
            yield data


def write_tfrecord(path, payloads):
    with open(path, "wb") as f:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + data + struct.pack("<I", masked_crc32c(data)))


# ------------------------------------------------------------------------------------------------ protos
def _ld(field, payload):
    return bytes([(field << 3) | 2]) + _varint(len(payload)) + payload


def serialize_tensor(a):
    """``tf.io.serialize_tensor``: TensorProto {dtype=1, tensor_shape=2 {dim=2 {size=1}}, tensor_content=4}."""
    a = np.asarray(a, order="C")
    dims = b"".join(_ld(2, b"\x08" + _varint(int(s))) for s in a.shape)
    return b"\x08" + _varint(_ENUM[a.dtype]) + _ld(2, dims) + _ld(4, a.astype(a.dtype.newbyteorder("<")).tobytes())


def parse_tensor(buf):
    """``tf.io.parse_tensor`` for numeric tensors (tensor_content, or the typed repeated fields small protos may use)."""
    p = _parse_proto(buf)
    dt = p.get(1, [0])[0]
    if dt not in _NP:
        raise ValueError(f"unsupported tensor dtype enum {dt}")
    shape = [_parse_proto(d).get(1, [0])[0] for d in _parse_proto(p[2][0]).get(2, [])] if 2 in p else []
    if 4 in p:
        return np.frombuffer(p[4][0], dtype=_NP[dt]).reshape(shape).copy()
    field = {_DT_FLOAT: 5, _DT_DOUBLE: 6, _DT_INT32: 7, _DT_INT64: 10}[dt]
    vals = []
    for v in p.get(field, []):
        if isinstance(v, bytes):                                  # packed
            if dt == _DT_FLOAT:
                vals += list(np.frombuffer(v, "<f4"))
            elif dt == _DT_DOUBLE:
                vals += list(np.frombuffer(v, "<f8"))
            else:
                pos = 0
                while pos < len(v):
                    from .tf_checkpoint import _read_varint
                    x, pos = _read_varint(v, pos)
                    vals.append(x - (1 << 64) if x >> 63 else x)
        else:
            vals.append(v)
    out = np.asarray(vals, dtype=_NP[dt])
    n = int(np.prod(shape)) if shape else 1
    if out.size == 1 and n > 1:
        out = np.full(n, out[0], dtype=_NP[dt])                   # TensorProto splat encoding
    return out.reshape(shape)


def _feature_bytes(v):
    return _ld(1, _ld(1, v))


def _feature_int64(v):
    return _ld(3, _ld(1, _varint(int(v) & 0xFFFFFFFFFFFFFFFF)))


def serialize_example(fid, text, mel, text_len, mel_len):
    """``TFRecordWriter.serialize_example`` (datasets/tf_record_utils.py:35-54): text int64 ids, mel float64."""
    feats = {"fid": _feature_bytes(fid.encode("utf-8")),
             "text": _feature_bytes(serialize_tensor(np.asarray(text, dtype=np.int64))),
             "mel": _feature_bytes(serialize_tensor(np.asarray(mel, dtype=np.float64))),
             "text_len": _feature_int64(text_len), "mel_len": _feature_int64(mel_len)}
    entries = b"".join(_ld(1, _ld(1, k.encode()) + _ld(2, v)) for k, v in sorted(feats.items()))
    return _ld(1, entries)


def parse_example_proto(buf):
    """{feature name: bytes | [floats] | [ints]} of a serialized ``tf.train.Example``."""
    from .tf_checkpoint import _read_varint
    out = {}
    feats = _parse_proto(_parse_proto(buf)[1][0])
    for entry in feats.get(1, []):
        e = _parse_proto(entry)
        name = e[1][0].decode()
        f = _parse_proto(e[2][0])
        if 1 in f:                                                # BytesList
            vals = _parse_proto(f[1][0]).get(1, [])
            out[name] = vals[0] if len(vals) == 1 else vals
        elif 2 in f:                                              # FloatList (packed or not)
            vals = []
            for v in _parse_proto(f[2][0]).get(1, []):
                vals += list(np.frombuffer(v, "<f4")) if isinstance(v, bytes) else [struct.unpack("<f", struct.pack("<I", v))[0]]
            out[name] = vals
        elif 3 in f:                                              # Int64List (packed or not)
            vals = []
            for v in _parse_proto(f[3][0]).get(1, []):
                if isinstance(v, bytes):
                    pos = 0
                    while pos < len(v):
                        x, pos = _read_varint(v, pos)
                        vals.append(x - (1 << 64) if x >> 63 else x)
                else:
                    vals.append(v - (1 << 64) if v >> 63 else v)
            out[name] = vals
    return out


def parse_example(buf, pad_factor=0):
    """``TFRecordWriter.parse_example`` (tf_record_utils.py:108-124): (fid, text int32 [T], mel float32 [M, num_mels],
    text_len, mel_len); ``pad_factor`` > 1 zero-pads the mel to a multiple of it (``pre_pad``, :93-106)."""
    ex = parse_example_proto(buf)
    text = parse_tensor(ex["text"]).astype(np.int32)
    mel = parse_tensor(ex["mel"]).astype(np.float32)
    if pad_factor > 1 and mel.shape[0] % pad_factor:
        mel = np.concatenate([mel, np.zeros((pad_factor - mel.shape[0] % pad_factor, mel.shape[1]), np.float32)])
    return ex["fid"].decode("utf-8"), text, mel, int(ex["text_len"][0]), int(ex["mel_len"][0])


# ------------------------------------------------------------------------------------------------ dataset
def iter_records(files, num_parallel_reads=1, verify_crc=True):
    """Record payloads in the order of ``tf.data.TFRecordDataset(files, num_parallel_reads=n)``: a deterministic
    interleave over ``n`` files at a time, one record per file per turn; exhausted files are replaced by the next one."""
    files = list(files)
    if num_parallel_reads <= 1:
        for f in files:
            yield from read_tfrecord(f, verify_crc)
        return
    pending = iter(files)
    active = []
    for _ in range(num_parallel_reads):
        f = next(pending, None)
        active.append(read_tfrecord(f, verify_crc) if f is not None else None)
    while any(a is not None for a in active):
        for i, it in enumerate(active):
            if it is None:
                continue
            rec = next(it, None)
            if rec is None:                                       # this slot moves on to the next file
                f = next(pending, None)
                active[i] = read_tfrecord(f, verify_crc) if f is not None else None
                if active[i] is not None:
                    rec = next(active[i], None)
            if rec is not None:
                yield rec


def padded_batch(items, num_mels, pin_memory=True):
    """Zero-padded batch (``padded_shapes=([], [None], [None, num_mels], [], [])``) as torch tensors:
    (fids, texts int32 [B, T_t], mels float32 [B, T_m, num_mels], text_len int32 [B], mel_len int32 [B])."""
    B = len(items)
    Tt = max(len(it[1]) for it in items)
    Tm = max(it[2].shape[0] for it in items)
    texts = torch.zeros(B, Tt, dtype=torch.int32)
    mels = torch.zeros(B, Tm, num_mels, dtype=torch.float32)
    t_len = torch.zeros(B, dtype=torch.int32)
    m_len = torch.zeros(B, dtype=torch.int32)
    for i, (_, text, mel, tl, ml) in enumerate(items):
        if mel.shape[1] != num_mels:
            raise ValueError(f"mel has {mel.shape[1]} bins, expected {num_mels}")
        texts[i, :len(text)] = torch.from_numpy(text)
        mels[i, :mel.shape[0]] = torch.from_numpy(mel)
        t_len[i], m_len[i] = tl, ml
    out = [texts, mels, t_len, m_len]
    if pin_memory and torch.cuda.is_available():
        out = [x.pin_memory() for x in out]
    return [it[0] for it in items], out[0], out[1], out[2], out[3]


def create_dataset(tfrecord_files, batch_size, num_mels, pad_factor=0, num_parallel_reads=1, shuffle=False,
                   shuffle_buffer=128, seed=1, pin_memory=True, shard=None):
    """``TFRecordWriter.create_dataset`` (tf_record_utils.py:126-142) as a generator of padded batches.
    ``shard=(rank, world)``: data-parallel training.  The stream is cut into GLOBAL batches of ``batch_size * world``
    utterances and rank r takes every world-th utterance of each global batch starting at r (the reference's own
    multi-worker rule ``utt_ids[rank::size]``, datasets/datasets.py:179-192, applied per global batch).  Every rank
    therefore yields the SAME number of batches of the SAME size -- each train_step issues collectives, unequal batch
    counts would deadlock at the end of an epoch; the tail that does not divide by ``world`` is dropped."""
    def batches():
        if shard is None:
            rank, world = 0, 1
        else:
            rank, world = int(shard[0]), int(shard[1])
            if not (0 <= rank < world):
                raise ValueError(f"shard=(rank {rank}, world {world})")
        cur = []
        for rec in iter_records(tfrecord_files, num_parallel_reads):
            cur.append(rec)
            if len(cur) == batch_size * world:
                yield padded_batch([parse_example(r, pad_factor) for r in cur[rank::world]], num_mels, pin_memory)
                cur = []
        if world > 1:
            cur = cur[:len(cur) // world * world]                 # same batch count and size on every rank
        if cur:
            yield padded_batch([parse_example(r, pad_factor) for r in cur[rank::world]], num_mels, pin_memory)   # final partial batch kept
    if not shuffle:
        yield from batches()
        return
    rng = np.random.default_rng(seed)
    buf = []
    for b in batches():                                           # Dataset.shuffle: fill a buffer, emit a random slot
        buf.append(b)
        if len(buf) > shuffle_buffer:
            j = int(rng.integers(len(buf)))
            buf[j], buf[-1] = buf[-1], buf[j]
            yield buf.pop()
    while buf:
        j = int(rng.integers(len(buf)))
        buf[j], buf[-1] = buf[-1], buf[j]
        yield buf.pop()
