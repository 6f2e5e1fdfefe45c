"""Reader / writer of TensorFlow checkpoint bundles (``<prefix>.index`` + ``<prefix>.data-0000N-of-0000M``), so that
DLC / DGP snapshots (``snapshot-step2-final--0``, resnet_v1_50.ckpt; reference: ``restorer.restore(sess, init_weights)`` at
src/deepgraphpose/models/fitdgp.py:689-720, src/deepgraphpose/models/eval.py:194-211, ``saver.save`` at fitdgp.py:830-839) load
into and save from the engine without TensorFlow.  SURVEY.md 8(f) rank 3.

Format (restated from the published TensorFlow sources, tensorflow==1.15: core/util/tensor_bundle/tensor_bundle.cc,
core/lib/io/{table,block,format}.cc -- a port of LevelDB's SSTable; core/protobuf/tensor_bundle.proto):

* ``.index`` is an SSTable.  Footer (last 48 bytes): metaindex BlockHandle, index BlockHandle (varint64 offset, varint64
  size each), zero padding to 40 bytes, magic 0xdb4775248b80fb57 (little endian).  A block = entries, then the uint32
  restart offsets, then uint32 num_restarts; on disk it is followed by a 1-byte compression type (0 none, 1 snappy) and a
  masked crc32c.  An entry = varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value.  The index
  block maps a separator key >= the last key of each data block to that block's BlockHandle.
* key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}; every other key is a variable name ->
  BundleEntryProto {1: dtype, 2: TensorShapeProto {2: dim {1: size}}, 3: shard_id, 4: offset, 5: size, 6: masked crc32c
  (fixed32) of the bytes, 7: slices (partitioned variables; not supported here)}.
* The data shards hold the raw little-endian tensor bytes.

PARITY UNPINNED: no TensorFlow-written bundle exists in the offline image or the reference tree to diff against; the tests
pin the reader against a hand-assembled table (prefix-compressed keys, multiple blocks, a snappy block) and the writer
against the reader.
"""
import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_CODES = {np.dtype(v): k for k, v in DTYPES.items()}
DT_BFLOAT16 = 14


# ----------------------------------------------------------------------------------------------------------- crc32c
def _make_table():
    poly = 0x82F63B78
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab.append(c)
    return np.array(tab, dtype=np.uint32)


_TAB = _make_table()


def crc32c(data):
    """CRC-32C (Castagnoli), as tensorflow/core/lib/hash/crc32c."""
    crc = 0xFFFFFFFF
    tab = _TAB
    for b in bytes(data):
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _crc32c_fast(data):
    """Same value through the native helper of libdgp_b200.so when it is built (a 100 MB checkpoint takes ~0.3 s instead of
    half a minute); pure Python otherwise."""
    raw = bytes(data)
    if len(raw) > 1 << 12:
        try:
            from . import _lib
            return int(_lib.load().dgp_crc32c(raw, len(raw)))
        except (ImportError, OSError, AttributeError):
            pass
    return crc32c(raw)


def mask_crc(crc):
    """crc32c::Mask: rotate right by 15 and add a constant (stored CRCs are masked)."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------------------- varints
def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("malformed varint")


def _put_varint(v):
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
    """Yield (field number, wire type, value) of one protobuf message (value: int for varint / fixed, bytes for length-delimited)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _field(field, wt, payload):
    head = _put_varint((field << 3) | wt)
    if wt == 0:
        return head + _put_varint(payload)
    if wt == 2:
        return head + _put_varint(len(payload)) + payload
    if wt == 5:
        return head + struct.pack("<I", payload)
    raise ValueError(wt)


# ----------------------------------------------------------------------------------------------------------- snappy
def snappy_uncompress(data):
    """Raw snappy block format (the only compression an SSTable block may carry)."""
    n, pos = _get_varint(data, 0)
    out = bytearray()
    while pos < len(data):
        tag = data[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(data[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += data[pos:pos + ln]
            pos += ln
        else:
            if kind == 1:
                ln = ((tag >> 2) & 7) + 4
                off = ((tag >> 5) << 8) | data[pos]
                pos += 1
            elif kind == 2:
                ln = (tag >> 2) + 1
                off = int.from_bytes(data[pos:pos + 2], "little")
                pos += 2
            else:
                ln = (tag >> 2) + 1
                off = int.from_bytes(data[pos:pos + 4], "little")
                pos += 4
            if off == 0 or off > len(out):
                raise ValueError("malformed snappy copy")
            for _ in range(ln):  # byte-wise: copies may overlap their own output
                out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ----------------------------------------------------------------------------------------------------------- SSTable
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError("truncated table block")
    body, ctype, stored = raw[:size], raw[size], struct.unpack_from("<I", raw, size + 1)[0]
    if verify and mask_crc(crc32c(raw[:size + 1])) != stored:
        raise ValueError("table block checksum mismatch")
    if ctype == 1:
        body = snappy_uncompress(body)
    elif ctype != 0:
        raise ValueError("unknown block compression %d" % ctype)
    return body


def _block_entries(block):
    if len(block) < 4:
        raise ValueError("malformed block")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        size = f.tell()
        if size < 48:
            raise ValueError("%s is too short to be a table" % path)
        f.seek(size - 48)
        footer = f.read(48)
        if struct.unpack_from("<Q", footer, 40)[0] != MAGIC:
            raise ValueError("%s: bad table magic (not a TensorFlow checkpoint index)" % path)
        pos = 0
        _, pos = _get_varint(footer, pos)
        _, pos = _get_varint(footer, pos)
        idx_off, pos = _get_varint(footer, pos)
        idx_size, pos = _get_varint(footer, pos)
        out = []
        for _, handle in _block_entries(_read_block(f, idx_off, idx_size, verify)):
            off, p2 = _get_varint(handle, 0)
            sz, _ = _get_varint(handle, p2)
            out.extend(_block_entries(_read_block(f, off, sz, verify)))
        return out


def _build_block(entries):
    """Block with every entry a restart point (shared = 0): valid for any LevelDB / TF reader."""
    body, restarts = bytearray(), []
    for k, v in entries:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(k)) + _put_varint(len(v)) + k + v
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def write_table(path, items, block_size=4096):
    """items: iterable of (key bytes, value bytes), keys strictly increasing."""
    items = list(items)
    for (a, _), (b, _) in zip(items[:-1], items[1:]):
        if not a < b:
            raise ValueError("table keys must be strictly increasing")
    with open(path, "wb") as f:
        def emit(block):
            off = f.tell()
            f.write(block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
            return off, len(block)

        index, cur, cur_size = [], [], 0
        for k, v in items:
            cur.append((k, v))
            cur_size += len(k) + len(v) + 8
            if cur_size >= block_size:
                off, sz = emit(_build_block(cur))
                index.append((cur[-1][0], _put_varint(off) + _put_varint(sz)))
                cur, cur_size = [], 0
        if cur or not index:
            off, sz = emit(_build_block(cur))
            index.append((cur[-1][0] if cur else b"", _put_varint(off) + _put_varint(sz)))
        meta_off, meta_sz = emit(_build_block([]))
        idx_off, idx_sz = emit(_build_block(index))
        footer = _put_varint(meta_off) + _put_varint(meta_sz) + _put_varint(idx_off) + _put_varint(idx_sz)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))


# ----------------------------------------------------------------------------------------------------------- bundle
def _parse_entry(value):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for field, wt, v in _parse_proto(value):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:
                    dim = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            dim = v3 if v3 < (1 << 63) else v3 - (1 << 64)
                    e["shape"].append(dim)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["slices"] += 1
    return e


def list_variables(prefix):
    """[(name, shape, numpy dtype or 'bfloat16')] of a checkpoint, like tf.train.list_variables."""
    out = []
    for k, v in read_table(prefix + ".index"):
        if k == b"":
            continue
        e = _parse_entry(v)
        out.append((k.decode(), tuple(e["shape"]), "bfloat16" if e["dtype"] == DT_BFLOAT16 else DTYPES.get(e["dtype"])))
    return out


def read_checkpoint(prefix, names=None, verify=False):
    """{variable name: ndarray} of the bundle ``prefix`` (``Saver.restore`` without TensorFlow).  ``names``: optional
    predicate or collection selecting variables; ``verify`` also checks the per-tensor crc32c (slow in pure Python)."""
    prefix = prefix[:-6] if prefix.endswith(".index") else prefix
    items = read_table(prefix + ".index")
    num_shards = 1
    for k, v in items:
        if k == b"":
            for field, _, val in _parse_proto(v):
                if field == 1:
                    num_shards = val
                elif field == 2 and val != 0:
                    raise ValueError("big-endian checkpoints are not supported")
    want = names if callable(names) or names is None else (lambda n, s=set(names): n in s)
    shards, out = {}, {}
    try:
        for k, v in items:
            if k == b"":
                continue
            name = k.decode()
            if want is not None and not want(name):
                continue
            e = _parse_entry(v)
            if e["slices"]:
                raise ValueError("%s: partitioned (sliced) variables are not supported" % name)
            if e["shard_id"] not in shards:
                shards[e["shard_id"]] = open("%s.data-%05d-of-%05d" % (prefix, e["shard_id"], num_shards), "rb")
            f = shards[e["shard_id"]]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError("%s: truncated data shard" % name)
            if verify and e["crc32c"] is not None and mask_crc(_crc32c_fast(raw)) != e["crc32c"]:
                raise ValueError("%s: tensor checksum mismatch" % name)
            if e["dtype"] == DT_BFLOAT16:
                arr = (np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16).view(np.float32)
            elif e["dtype"] in DTYPES:
                arr = np.frombuffer(raw, dtype=np.dtype(DTYPES[e["dtype"]]).newbyteorder("<"))
            else:
                raise ValueError("%s: unsupported dtype enum %d" % (name, e["dtype"]))
            out[name] = arr.reshape(e["shape"]).copy()
    finally:
        for f in shards.values():
            f.close()
    return out


def write_checkpoint(prefix, variables, with_crc=True):
    """Write {name: ndarray} as a single-shard bundle readable by ``tf.train.Saver`` / ``tf.train.load_checkpoint``."""
    prefix = prefix[:-6] if prefix.endswith(".index") else prefix
    header = _field(1, 0, 1) + _field(2, 0, 0) + _field(3, 2, _field(1, 0, 1))
    items = [(b"", header)]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(variables, key=lambda s: s.encode()):
            a = np.asarray(variables[name])
            if not a.flags.c_contiguous:   # (np.ascontiguousarray would turn a scalar such as global_step into shape (1,))
                a = a.copy(order="C")
            if a.dtype not in DTYPE_CODES:
                raise ValueError("%s: dtype %s has no TensorFlow enum here" % (name, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            shape = b"".join(_field(2, 2, _field(1, 0, int(d))) for d in a.shape)
            entry = _field(1, 0, DTYPE_CODES[a.dtype]) + _field(2, 2, shape) + _field(3, 0, 0) + _field(4, 0, offset) + \
                _field(5, 0, len(raw))
            if with_crc:
                entry += _field(6, 5, mask_crc(_crc32c_fast(raw)))
            items.append((name.encode(), entry))
            offset += len(raw)
    write_table(prefix + ".index", items)


def model_variables(variables):
    """The variables the reference's restorer covers (fitdgp.py:689-696: scopes resnet*, pose/part_pred, pose/locref_pred),
    i.e. the checkpoint minus optimizer slots (``.../Momentum``), ``global_step`` and the like."""
    keep = {}
    for k, v in variables.items():
        if k.endswith("/Momentum") or "/Momentum_" in k or k.endswith("/Adam") or k.endswith("/Adam_1"):
            continue
        if k.startswith("resnet_v1_") or k.startswith("pose/part_pred") or k.startswith("pose/locref_pred"):
            keep[k] = v
    return keep
