"""Reader (and writer) for TensorFlow's V2 checkpoint files -- ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` --
without TensorFlow (what ``saver.restore(sess, latest_ckpt)`` reads in the reference,
train.py:49-58 / test.py:40-48, and what download_weights.sh:4 delivers for the released Nancy weights).

Format, restated from TensorFlow's ``core/util/tensor_bundle`` and ``core/lib/io/table`` (the LevelDB table format):

  ``.index`` is an SSTable.  Data / index blocks hold prefix-compressed entries
      [shared varint32][non_shared varint32][value_len varint32][key suffix][value]
  followed by the restart array (uint32 offsets) and its length (uint32); every block is followed by a 5-byte trailer
  (1 byte compression type: 0 = none, 1 = snappy; 4 bytes masked crc32c).  The 48-byte footer holds the BlockHandles
  (varint64 offset, varint64 size) of the metaindex and index blocks, zero padding, and the magic 0xdb4775248b80fb57.
  The index block maps a separator key to the BlockHandle of each data block.  Keys are tensor names; the empty key ""
  carries a BundleHeaderProto, every other value a BundleEntryProto:
      1 dtype (enum)   2 shape (TensorShapeProto: repeated 2 dim { 1 size })   3 shard_id   4 offset   5 size   6 crc32c (fixed32)
  ``.data-0000N-of-0000M`` is the concatenation of the raw little-endian tensor bytes at those offsets.

The tensor-bundle writer of TF 1.x does not compress the index (``table::kNoCompression``); a snappy block raises.
PINNING: no TensorFlow-written checkpoint is reachable in this environment, so this reader is checked against (a) the
writer below, which follows the same specification, and (b) hand-assembled byte strings of the block / varint / protobuf
layers (tests/test_tf_checkpoint.py).  It has NOT met a file written by TensorFlow itself.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_ENUM = {np.dtype(v): k for k, v in DTYPES.items()}


class BundleError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# varints / crc32c / protobuf wire format
# ---------------------------------------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise BundleError("varint too long")


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


_CRC_TABLE = None


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), the checksum of LevelDB blocks and of tensor data"""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    c = crc ^ 0xffffffff
    for b in bytes(data):
        c = _CRC_TABLE[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def mask_crc(c):
    """leveldb's crc masking: rotate right by 15 bits and add a constant"""
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


def _parse_proto(buf):
    """flat protobuf message -> {field number: [values]} (varint -> int, 64/32-bit -> int, length-delimited -> bytes)"""
    out = {}
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n]); pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise BundleError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


# ---------------------------------------------------------------------------------------------------------------------
# SSTable
# ---------------------------------------------------------------------------------------------------------------------
def _read_block(f, offset, size, verify):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise BundleError("truncated table block")
    block, ctype, crc = raw[:size], raw[size], struct.unpack("<I", raw[size + 1:])[0]
    if verify and mask_crc(crc32c(raw[:size + 1])) != crc:
        raise BundleError(f"table block at {offset}: crc mismatch")
    if ctype == 1:
        raise BundleError("snappy-compressed table block (TF's bundle writer does not compress; unsupported here)")
    if ctype != 0:
        raise BundleError(f"unknown block compression type {ctype}")
    return block


def _block_entries(block):
    if len(block) < 4:
        raise BundleError("table block too short")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise BundleError("bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise BundleError("corrupt table entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(index_path, verify_crc=True):
    """.index -> (header dict, {tensor name: entry dict})"""
    with open(index_path, "rb") as f:
        f.seek(0, os.SEEK_END)
        size = f.tell()
        if size < FOOTER_LEN:
            raise BundleError("index file shorter than a table footer")
        f.seek(size - FOOTER_LEN)
        footer = f.read(FOOTER_LEN)
        if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
            raise BundleError("not an SSTable (bad magic): is this a V1 checkpoint?")
        pos = 0
        _, pos = _get_varint(footer, pos); _, pos = _get_varint(footer, pos)          # metaindex handle
        ioff, pos = _get_varint(footer, pos); isz, pos = _get_varint(footer, pos)      # index handle
        header, entries = None, {}
        for _, handle in _block_entries(_read_block(f, ioff, isz, verify_crc)):
            boff, p = _get_varint(handle, 0)
            bsz, p = _get_varint(handle, p)
            for key, value in _block_entries(_read_block(f, boff, bsz, verify_crc)):
                msg = _parse_proto(value)
                if key == b"":
                    header = {"num_shards": msg.get(1, [1])[0], "endianness": msg.get(2, [0])[0]}
                    continue
                shape = []
                for sp in msg.get(2, []):
                    for dim in _parse_proto(sp).get(2, []):
                        shape.append(_signed64(_parse_proto(dim).get(1, [0])[0]))
                entries[key.decode("utf-8")] = {
                    "dtype": msg.get(1, [0])[0], "shape": tuple(shape), "shard_id": msg.get(3, [0])[0],
                    "offset": msg.get(4, [0])[0], "size": msg.get(5, [0])[0], "crc32c": msg.get(6, [None])[0],
                    "sliced": bool(msg.get(7))}
    if header is None:
        raise BundleError("bundle header entry (empty key) missing")
    if header["endianness"] != 0:
        raise BundleError("big-endian bundle")
    return header, entries


def read_bundle(prefix, names=None, verify_crc=True):
    """{variable name: numpy array} of the checkpoint ``prefix`` (the string passed to saver.save / saver.restore)."""
    header, entries = read_index(prefix + ".index", verify_crc)
    out = {}
    files = {}
    try:
        for name, e in entries.items():
            if names is not None and name not in names:
                continue
            if e["sliced"]:
                raise BundleError(f"{name}: partitioned (sliced) variables are not supported")
            if e["dtype"] not in DTYPES:
                raise BundleError(f"{name}: unsupported dtype enum {e['dtype']}")
            dt = np.dtype(DTYPES[e["dtype"]])
            count = int(np.prod(e["shape"])) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise BundleError(f"{name}: size {e['size']} does not match shape {e['shape']} of {dt}")
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open(f"{prefix}.data-{sid:05d}-of-{header['num_shards']:05d}", "rb")
            fh = files[sid]
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
            if len(raw) != e["size"]:
                raise BundleError(f"{name}: data file truncated")
            if verify_crc and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
                raise BundleError(f"{name}: tensor crc mismatch")
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for fh in files.values():
            fh.close()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# writer (same specification; used for the round-trip tests and to hand weights to a TensorFlow installation)
# ---------------------------------------------------------------------------------------------------------------------
def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _entry_proto(dtype_enum, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s) & ((1 << 64) - 1))))) for s in shape)
    msg = _field(1, 0, _put_varint(dtype_enum)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _put_varint(offset))
    msg += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return msg


def _build_block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix, tensors, block_size=4096):
    """write {name: array} as ``prefix.index`` + ``prefix.data-00000-of-00001`` (one shard, uncompressed index)"""
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    items = [(b"", _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(1)))))]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as df:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])
            if a.ndim and not a.flags.c_contiguous:          # (ascontiguousarray would turn a scalar into shape (1,))
                a = np.ascontiguousarray(a)
            if a.dtype not in DTYPE_ENUM:
                raise BundleError(f"{name}: dtype {a.dtype} has no TensorFlow enum here")
            raw = a.tobytes()
            df.write(raw)
            items.append((name.encode("utf-8"), _entry_proto(DTYPE_ENUM[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    with open(prefix + ".index", "wb") as f:
        index_items, cur, cur_bytes = [], [], 0

        def flush():
            nonlocal cur, cur_bytes
            if not cur:
                return
            block = _build_block(cur)
            off = f.tell()
            f.write(block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
            index_items.append((cur[-1][0], _put_varint(off) + _put_varint(len(block))))     # separator = last key of the block
            cur, cur_bytes = [], 0
        for k, v in items:
            cur.append((k, v))
            cur_bytes += len(k) + len(v) + 6
            if cur_bytes >= block_size:
                flush()
        flush()
        meta = _build_block([])
        moff = f.tell()
        f.write(meta + b"\x00" + struct.pack("<I", mask_crc(crc32c(meta + b"\x00"))))
        idx = _build_block(index_items, restart_interval=1)
        ioff = f.tell()
        f.write(idx + b"\x00" + struct.pack("<I", mask_crc(crc32c(idx + b"\x00"))))
        footer = _put_varint(moff) + _put_varint(len(meta)) + _put_varint(ioff) + _put_varint(len(idx))
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
        f.write(footer)
    return prefix


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint without the `checkpoint` state file: the .index with the largest step suffix"""
    import glob
    import re
    best = None
    for p in glob.glob(os.path.join(directory, "*.index")):
        m = re.search(r"-(\d+)\.index$", p)
        step = int(m.group(1)) if m else -1
        if best is None or step > best[0]:
            best = (step, p[:-len(".index")])
    return best[1] if best else None
