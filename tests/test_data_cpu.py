"""TFRecord / tf.train.Example input pipeline without TensorFlow (vaenar_tts_b200/data.py):
known-answer encodings, corruption detection, the padded-batch contract of datasets/tf_record_utils.py:126-142.
PARITY UNPINNED against files written by TensorFlow (none can be produced in this image); CPU only."""
import struct

import numpy as np
import pytest
import torch

from vaenar_tts_b200 import data as D
from vaenar_tts_b200.tf_checkpoint import masked_crc32c


def test_hand_encoded_example():
    # Example{features{feature{"a": int64_list{value:[5]}}}} written out by hand from the protobuf wire format
    feature = bytes([0x1A, 0x03, 0x0A, 0x01, 0x05])
    entry = bytes([0x0A, 0x01, 0x61, 0x12, 0x05]) + feature
    features = bytes([0x0A, len(entry)]) + entry
    example = bytes([0x0A, len(features)]) + features
    assert D.parse_example_proto(example) == {"a": [5]}
    # unpacked Int64List (older writers) and a negative value
    feature = bytes([0x1A, 0x0B, 0x08]) + bytes([0xFF] * 9 + [0x01])
    entry = bytes([0x0A, 0x01, 0x62, 0x12, len(feature)]) + feature
    features = bytes([0x0A, len(entry)]) + entry
    assert D.parse_example_proto(bytes([0x0A, len(features)]) + features) == {"b": [-1]}
    # FloatList packed
    fl = struct.pack("<2f", 1.5, -2.0)
    feature = bytes([0x12, 2 + len(fl), 0x0A, len(fl)]) + fl
    entry = bytes([0x0A, 0x01, 0x63, 0x12, len(feature)]) + feature
    features = bytes([0x0A, len(entry)]) + entry
    assert D.parse_example_proto(bytes([0x0A, len(features)]) + features) == {"c": [1.5, -2.0]}


def test_tensor_proto_roundtrip_and_typed_fields():
    for a in (np.arange(12, dtype=np.int64).reshape(3, 4), np.random.default_rng(0).random((5, 80)),
              np.asarray(7, dtype=np.int64), np.zeros((0, 80))):
        b = D.parse_tensor(D.serialize_tensor(a))
        assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)
    # TensorProto with int64_val (field 10) instead of tensor_content: dtype=DT_INT64, shape [3], packed values 1,2,3
    proto = bytes([0x08, 0x09, 0x12, 0x04, 0x12, 0x02, 0x08, 0x03, 0x52, 0x03, 0x01, 0x02, 0x03])
    assert D.parse_tensor(proto).tolist() == [1, 2, 3]


def test_record_framing_and_corruption(tmp_path):
    path = str(tmp_path / "t.tfrecords")
    payloads = [b"", b"abc", bytes(range(256)) * 20]
    D.write_tfrecord(path, payloads)
    raw = open(path, "rb").read()
    assert raw[:8] == struct.pack("<Q", 0) and struct.unpack("<I", raw[8:12])[0] == masked_crc32c(raw[:8])
    assert list(D.read_tfrecord(path)) == payloads
    bad = bytearray(raw)
    bad[-10] ^= 0x01
    open(path, "wb").write(bad)
    with pytest.raises(ValueError):
        list(D.read_tfrecord(path))
    assert len(list(D.read_tfrecord(path, verify_crc=False))) == 3


def _utt(i, rng):
    tl, ml = int(rng.integers(3, 12)), int(rng.integers(10, 40))
    return f"LJ{i:03d}", rng.integers(1, 43, tl), rng.random((ml, 80)), tl, ml


def test_create_dataset_contract(tmp_path):
    """dtypes / padding / order of create_dataset: int32 texts + lengths, float32 mels, zero padding to the longest item,
    final partial batch kept, deterministic interleave over files, pad_factor, sharding rule utt_ids[rank::size]"""
    rng = np.random.default_rng(1)
    utts = [_utt(i, rng) for i in range(11)]
    files = []
    for k in range(2):                                            # train-0 / train-1 like TFRecordWriter.write
        p = str(tmp_path / f"train-{k}.tfrecords")
        D.write_tfrecord(p, [D.serialize_example(*u) for u in utts[k::2]])
        files.append(p)
    batches = list(D.create_dataset(files, batch_size=4, num_mels=80, pin_memory=False))
    assert [len(b[0]) for b in batches] == [4, 4, 3]
    seq = utts[0::2] + utts[1::2]                                  # files are read one after the other
    fids = [f for b in batches for f in b[0]]
    assert fids == [u[0] for u in seq]
    fid, texts, mels, t_len, m_len = batches[0]
    assert texts.dtype == torch.int32 and mels.dtype == torch.float32 and t_len.dtype == torch.int32 and m_len.dtype == torch.int32
    assert texts.shape == (4, int(t_len.max())) and mels.shape == (4, int(m_len.max()), 80)
    for i, u in enumerate(seq[:4]):
        assert texts[i, :u[3]].tolist() == list(u[1]) and int(texts[i, u[3]:].abs().sum()) == 0
        assert torch.allclose(mels[i, :u[4]], torch.from_numpy(u[2]).float()) and float(mels[i, u[4]:].abs().sum()) == 0.0
    # num_parallel_reads=2: one record per file per turn
    inter = [f for b in D.create_dataset(files, 4, 80, num_parallel_reads=2, pin_memory=False) for f in b[0]]
    assert inter == [u[0] for u in utts]
    # pad_factor: mel length padded to a multiple, mel_len untouched
    b0 = next(D.create_dataset(files, 2, 80, pad_factor=8, pin_memory=False))
    assert b0[2].shape[1] % 8 == 0 and b0[2].shape[1] >= int(b0[4].max())
    # shuffle permutes whole batches; sharding is a partition
    shuf = [tuple(b[0]) for b in D.create_dataset(files, 4, 80, shuffle=True, shuffle_buffer=2, seed=3, pin_memory=False)]
    assert sorted(shuf) == sorted(tuple(b[0]) for b in batches)
    parts = [[f for b in D.create_dataset(files, 4, 80, pin_memory=False, shard=(r, 2)) for f in b[0]] for r in range(2)]
    assert not set(parts[0]) & set(parts[1])
    assert parts[0] == fids[0:10:2] and parts[1] == fids[1:10:2]         # 11 utterances: the odd one out is dropped


@pytest.mark.parametrize("n_utt,world,bs", [(11, 2, 4), (13, 4, 2), (16, 4, 2), (7, 8, 1), (33, 3, 4)])
def test_sharded_dataset_same_batch_count_on_every_rank(tmp_path, n_utt, world, bs):
    """Every train_step issues collectives: all ranks must see the same number of batches (and batch sizes) even when the
    record count does not divide by world * batch_size -- otherwise the rank with the extra batch deadlocks."""
    rng = np.random.default_rng(5)
    utts = [_utt(i, rng) for i in range(n_utt)]
    p = str(tmp_path / "train-0.tfrecords")
    D.write_tfrecord(p, [D.serialize_example(*u) for u in utts])
    per_rank = [[tuple(b[0]) for b in D.create_dataset([p], bs, 80, pin_memory=False, shard=(r, world))] for r in range(world)]
    sizes = [[len(b) for b in pr] for pr in per_rank]
    assert all(s == sizes[0] for s in sizes), sizes
    seen = [f for pr in per_rank for b in pr for f in b]
    assert len(seen) == len(set(seen)) == n_utt // world * world if n_utt >= world else len(seen) == 0
    for r in range(world):
        flat = [f for b in per_rank[r] for f in b]
        assert flat == [u[0] for u in utts[:n_utt // world * world]][r::world] or sizes[0] == []
