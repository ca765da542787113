"""TF tensor-bundle checkpoint reader / writer (vaenar_tts_b200/tf_checkpoint.py): known-answer
vectors of the format primitives and write -> read round trips.  PARITY UNPINNED against real TensorFlow files (none can
be produced in this image); CPU only."""
import os
import struct

import numpy as np
import pytest
import torch

from vaenar_tts_b200 import tf_checkpoint as C


def test_crc32c_known_answers():
    # RFC 3720 / iSCSI test vectors
    assert C.crc32c(b"123456789") == 0xE3069283
    assert C.crc32c(bytes(32)) == 0x8A9136AA
    assert C.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert C.crc32c(bytes(range(32))) == 0x46DD794E
    big = bytes(range(256)) * 64                      # > 4096 bytes: the native (C-ABI) implementation
    py = 0
    c = 0 ^ 0xFFFFFFFF
    for b in big:
        c = C._TBL_LIST[(c ^ b) & 0xFF] ^ (c >> 8)
    assert C.crc32c(big) == (c ^ 0xFFFFFFFF)
    # LevelDB's mask: rotate right 15, add 0xa282ead8
    v = C.crc32c(b"foo")
    assert C.masked_crc32c(b"foo") == ((((v >> 15) | (v << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_varint_and_proto():
    for v in (0, 1, 127, 128, 300, 2 ** 31 - 1, 2 ** 40 + 5):
        enc = C._varint(v)
        assert C._read_varint(enc, 0) == (v, len(enc))
    assert C._varint(300) == b"\xac\x02"                                   # protobuf documentation example
    e = C._parse_proto(C._entry_proto(1, (3, 5), 1024, 60, 0xDEADBEEF))
    assert e[1] == [1] and e[4] == [1024] and e[5] == [60] and e[6] == [0xDEADBEEF]
    dims = [C._parse_proto(d)[1][0] for d in C._parse_proto(e[2][0])[2]]
    assert dims == [3, 5]


def test_table_roundtrip_and_footer(tmp_path):
    items = [(("key%05d" % i).encode(), os.urandom(i % 37)) for i in range(500)] + [(b"", b"hdr")]
    path = str(tmp_path / "t.index")
    C.write_table(path, items)
    raw = open(path, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 48
    assert C.read_table(path) == sorted(items)
    corrupted = bytearray(raw)
    corrupted[10] ^= 0x40
    open(path, "wb").write(corrupted)
    with pytest.raises(ValueError):
        C.read_table(path)


def test_name_mapping():
    k = "model/prior/glow/3/2/net/attentions/1/ffn/dense2/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    n = "prior.glow.3.affine_coupling.net.attentions.1.ffn.dense2.kernel"
    assert C.tf_key_to_name(k) == n and C.name_to_tf_key(n) == k
    assert C.tf_key_to_name("model/prior/glow/0/0/log_scale/.ATTRIBUTES/VARIABLE_VALUE") == "prior.glow.0.actnorm.log_scale"
    assert C.tf_key_to_name("model/prior/glow/5/1/weight/.ATTRIBUTES/VARIABLE_VALUE") == "prior.glow.5.linear.weight"
    assert C.tf_key_to_name("model/text_encoder/self_attentions/2/attention/query_layer/kernel/.ATTRIBUTES/VARIABLE_VALUE") == \
        "text_encoder.self_attentions.2.attention.query_layer.kernel"
    assert C.tf_key_to_name("optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE") is None
    assert C.tf_key_to_name("model/decoder/pre_projection/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE") is None
    assert C.tf_key_to_name("_CHECKPOINTABLE_OBJECT_GRAPH") is None


def test_state_dict_roundtrip_through_bundle(tmp_path):
    """every parameter name of the manifest survives save -> load, bit for bit, incl. the (actnorm, linear, coupling) tuples"""
    import __graft_entry__ as g
    g.build()
    from vaenar_tts_b200 import VAENAR, LJHPS
    m = VAENAR(LJHPS, device="cpu", seed=3)
    sd = m.state_dict()
    prefix = str(tmp_path / "ckpt-7")
    C.save_tf_checkpoint(prefix, sd, step=7)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    raw = C.read_bundle(prefix)
    assert int(np.asarray(raw["step/.ATTRIBUTES/VARIABLE_VALUE"]).reshape(-1)[0]) == 7
    back = C.load_tf_checkpoint(prefix)
    assert set(back) == set(sd)
    for k, v in sd.items():
        want = () if k.endswith("pos_weight") else tuple(v.shape)      # TF scalars (tf.Variable(1.0)) have shape []
        assert back[k].shape == want and np.array_equal(back[k].reshape(v.shape), v.numpy()), k
    m2 = VAENAR(LJHPS, device="cpu", seed=4)
    m2.load_state_dict(back)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    # a flipped data byte is caught by the per-tensor CRC32C
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[1234] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    with pytest.raises(ValueError):
        C.read_bundle(prefix)


def test_optimizer_state_and_step_survive_a_checkpoint(tmp_path):
    """train.py:246 checkpoints (step, optimizer, model): the Adam moments and the optimizer iteration must come back, so that
    a resumed run continues the bias correction instead of restarting at t = 1 with zero moments."""
    import __graft_entry__ as g
    g.build()
    from vaenar_tts_b200 import VAENAR, LJHPS
    m = VAENAR(LJHPS, device="cpu", seed=3)
    m._ensure_adam_state()
    gen = torch.Generator().manual_seed(0)
    m._adam_m.copy_(torch.randn(m._adam_m.shape, generator=gen) * m._trainable_mask.float())
    m._adam_v.copy_(torch.rand(m._adam_v.shape, generator=gen) * m._trainable_mask.float())
    m._opt_step = 1234
    prefix = str(tmp_path / "ckpt-3")
    m.save_tf_checkpoint(prefix, step=3)
    raw = C.read_bundle(prefix)
    assert raw["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"].shape == () and int(raw["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"]) == 1234
    assert raw["model/text_encoder/pos_weight/.ATTRIBUTES/VARIABLE_VALUE"].shape == ()
    key = "model/decoder/pre_projection/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"
    assert key in raw and raw[key].shape == (128, 256)
    m2 = VAENAR(LJHPS, device="cpu", seed=9)
    m2.load_tf_checkpoint(prefix)
    assert m2._opt_step == 1234
    assert torch.equal(m2._adam_m, m._adam_m) and torch.equal(m2._adam_v, m._adam_v)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
