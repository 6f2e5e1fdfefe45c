"""TensorFlow checkpoint-bundle reader / writer (deepgraphpose_b200/tf_checkpoint.py), CPU only.  No TensorFlow-written file is
available offline, so the reader is pinned against a table assembled by hand from the published format (prefix-compressed
keys, several data blocks, restart arrays, a snappy block), CRC-32C against its published check value, and the writer against
the reader."""
import os
import struct

import numpy as np
import pytest

from deepgraphpose_b200 import tf_checkpoint as ck


def test_crc32c_known_answers(lib_built):
    assert ck.crc32c(b"123456789") == 0xE3069283          # the CRC-32C check value
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(bytes(32)) == 0x8A9136AA              # RFC 3720 B.4: 32 zero bytes
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43     # RFC 3720 B.4: 32 bytes of 0xff
    big = np.random.default_rng(0).integers(0, 256, 100000, dtype=np.uint8).tobytes()
    assert ck._crc32c_fast(big) == ck.crc32c(big)          # native helper == pure Python
    assert ck.mask_crc(0) == 0xA282EAD8


def test_snappy_decoder():
    # literal "abcd", then copy(offset 4, len 8) overlapping its own output, then literal "XY"
    comp = bytes([14]) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([(2 - 1) << 2]) + b"XY"
    assert ck.snappy_uncompress(comp) == b"abcdabcdabcdXY"
    with pytest.raises(ValueError):
        ck.snappy_uncompress(bytes([3, 1, 5]))             # copy before any output


def _block(entries, restart_interval=2):
    """LevelDB block with shared-prefix compression (what TensorFlow's table builder emits)."""
    body, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(k), len(prev)) and k[shared] == prev[shared]:
                shared += 1
        body += ck._put_varint(shared) + ck._put_varint(len(k) - shared) + ck._put_varint(len(v)) + k[shared:] + v
        prev = k
    for r in restarts:
        body += struct.pack("<I", r)
    return bytes(body + struct.pack("<I", len(restarts)))


def test_reader_on_hand_assembled_table(tmp_path):
    """Two data blocks with prefix-compressed keys (one stored snappy-'compressed' as a single literal), an index block with
    separator keys that are NOT equal to the last key of the block, a non-empty-looking metaindex, footer + magic."""
    kv = [(b"", b"hdr"), (b"resnet_v1_50/conv1/BatchNorm/beta", b"v1"), (b"resnet_v1_50/conv1/BatchNorm/gamma", b"v2"),
          (b"resnet_v1_50/conv1/weights", b"v3"), (b"resnet_v1_50/conv2/weights", b"v4"), (b"z", b"v5")]
    path = str(tmp_path / "t.index")
    with open(path, "wb") as f:
        def emit(block, ctype=0):
            off = f.tell()
            payload = block
            if ctype == 1:   # snappy: uncompressed length + one literal covering the whole block
                ln = len(block) - 1
                payload = ck._put_varint(len(block)) + bytes([61 << 2]) + struct.pack("<H", ln) + block
            f.write(payload + bytes([ctype]) + struct.pack("<I", ck.mask_crc(ck.crc32c(payload + bytes([ctype])))))
            return ck._put_varint(off) + ck._put_varint(len(payload))
        h1 = emit(_block(kv[:4]))
        h2 = emit(_block(kv[4:]), ctype=1)
        meta = emit(_block([]))
        index = emit(_block([(b"resnet_v1_50/conv1/x", h1), (b"zz", h2)], restart_interval=1))
        footer = meta + index
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", ck.MAGIC))
    assert ck.read_table(path) == kv
    bad = open(path, "rb").read()
    open(path, "wb").write(bad[:10] + bytes([bad[10] ^ 1]) + bad[11:])
    with pytest.raises(ValueError):
        ck.read_table(path)                                  # block checksum
    open(path, "wb").write(bad[:-1] + b"\x00")
    with pytest.raises(ValueError):
        ck.read_table(path)                                  # magic


def test_bundle_entry_proto_parsing():
    # dtype DT_FLOAT, shape [3, 3, 64, 256], shard 0, offset 1234, size 589824, crc 0xdeadbeef
    shape = b"".join(ck._field(2, 2, ck._field(1, 0, d)) for d in (3, 3, 64, 256))
    e = ck._parse_entry(ck._field(1, 0, 1) + ck._field(2, 2, shape) + ck._field(4, 0, 1234) + ck._field(5, 0, 589824) +
                        ck._field(6, 5, 0xDEADBEEF))
    assert (e["dtype"], e["shape"], e["shard_id"], e["offset"], e["size"], e["crc32c"]) == (1, [3, 3, 64, 256], 0, 1234, 589824, 0xDEADBEEF)
    assert ck._parse_entry(ck._field(1, 0, 3))["shape"] == []    # scalar (global_step)


def test_write_read_round_trip(tmp_path, lib_built):
    from deepgraphpose_b200 import synthetic
    W = synthetic.make_weights(4, seed=1)
    W["global_step"] = np.array(1234, dtype=np.int64)
    W["resnet_v1_50/conv1/weights/Momentum"] = np.zeros((7, 7, 3, 64), np.float32)
    prefix = str(tmp_path / "snapshot-step2-final--0")
    ck.write_checkpoint(prefix, W)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    got = ck.read_checkpoint(prefix, verify=True)
    assert sorted(got) == sorted(W)
    for k in W:
        assert got[k].dtype == W[k].dtype and got[k].shape == W[k].shape and np.array_equal(got[k], W[k]), k
    names = [n for n, _, _ in ck.list_variables(prefix)]
    assert names == sorted(W, key=lambda s: s.encode()) and len(names) > 260    # many index blocks
    model = ck.model_variables(got)
    assert "global_step" not in model and not any(k.endswith("/Momentum") for k in model) and len(model) == len(W) - 2
    sub = ck.read_checkpoint(prefix + ".index", names=["pose/part_pred/block4/biases"])
    assert list(sub) == ["pose/part_pred/block4/biases"]
    # corrupt one tensor byte: the per-tensor checksum catches it
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(100)
        b = f.read(1)
        f.seek(100)
        f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(ValueError):
        ck.read_checkpoint(prefix, verify=True)


def test_load_variables_accepts_a_checkpoint_prefix(tmp_path, lib_built):
    from deepgraphpose_b200 import synthetic
    from deepgraphpose_b200.eval import load_variables
    W = synthetic.make_weights(4, seed=2, location_refinement=False)
    prefix = str(tmp_path / "snapshot-5")
    ck.write_checkpoint(prefix, dict(W, global_step=np.array(5, dtype=np.int64)), with_crc=False)
    got = load_variables(prefix, 4, False)
    assert sorted(got) == sorted(W) and all(np.array_equal(got[k], W[k]) for k in W)
