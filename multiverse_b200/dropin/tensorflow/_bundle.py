# coding=utf-8
"""Reader (and a writer, for tests) of TensorFlow's checkpoint "tensor bundle" files, in pure Python
(SURVEY.md §8 row f-2): `<prefix>.index` + `<prefix>.data-00000-of-0000N`, the format
`tf.train.Saver` writes since TF 1.0 and the one the released Multiverse models ship in
(TESTING.md "download the pretrained models"; restored by code/pred_utils.py:186-204).

Format (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc}, lib/io/table*.cc - a port of LevelDB's table):
  * `.index` is an immutable sorted string table: data blocks, a metaindex block, an index block and a
    48-byte footer {metaindex handle, index handle, padding, magic 0xdb4775248b80fb57}.  A block is a run of
    prefix-compressed entries {shared varint32, unshared varint32, value_len varint32, key suffix, value},
    a uint32 restart array and its length; each block is followed by {compression type u8, masked crc32c u32}.
  * key "" -> BundleHeaderProto {num_shards=1, endianness=2, version=3};
    key <variable name> -> BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6, slices=7}.
  * `.data-XXXXX-of-YYYYY` holds the raw little-endian tensor bytes at [offset, offset+size).

Not validated against a file produced by TensorFlow itself: no TensorFlow and no checkpoint exist in the
build container, so the writer below (same specification) is the only producer the reader was tested with
(tests/test_dropin_cpu.py).  Snappy-compressed blocks (never written by tf.train.Saver) are refused loudly.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
          10: np.bool_, 14: None, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_ENUM = {np.dtype(v): k for k, v in DTYPES.items() if v is not None}


# --------------------------------------------------------------------------- crc32c (Castagnoli), masked as LevelDB
def _make_table():
  t = []
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    t.append(c)
  return t


_CRC_TABLE = _make_table()


def crc32c(data, crc=0):
  crc ^= 0xFFFFFFFF
  for b in data:
    crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
  return crc ^ 0xFFFFFFFF


def mask_crc(crc):
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# --------------------------------------------------------------------------- varints / minimal protobuf
def _get_varint(buf, pos):
  out, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    out |= (b & 0x7F) << shift
    if not b & 0x80:
      return out, pos
    shift += 7


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
  """-> list of (field number, wire type, value); value is int (varint / fixed) or bytes (length-delimited)."""
  pos, out = 0, []
  while pos < len(buf):
    tag, pos = _get_varint(buf, pos)
    field, wt = tag >> 3, tag & 7
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
      raise ValueError("unsupported protobuf wire type %d" % wt)
    out.append((field, wt, v))
  return out


def _signed64(v):
  return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
  dims = []
  for field, _, v in _parse_proto(buf):
    if field == 2:                                  # Dim { int64 size = 1; string name = 2 }
      size = 0
      for f2, _, v2 in _parse_proto(v):
        if f2 == 1:
          size = _signed64(v2)
      dims.append(size)
    elif field == 3 and v:
      raise ValueError("tensor of unknown rank in checkpoint")
  return tuple(dims)


def _parse_entry(buf):
  e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, slices=False)
  for field, _, v in _parse_proto(buf):
    if field == 1: e["dtype"] = v
    elif field == 2: e["shape"] = _parse_shape(v)
    elif field == 3: e["shard_id"] = v
    elif field == 4: e["offset"] = v
    elif field == 5: e["size"] = v
    elif field == 6: e["crc32c"] = v
    elif field == 7: e["slices"] = True
  return e


# --------------------------------------------------------------------------- table reader
def _read_block(data, offset, size, verify=True):
  block = data[offset:offset + size]
  ctype = data[offset + size]
  if verify:
    want = struct.unpack_from("<I", data, offset + size + 1)[0]
    got = mask_crc(crc32c(data[offset:offset + size + 1]))
    if want != got:
      raise IOError("corrupt checkpoint index: block crc mismatch at offset %d" % offset)
  if ctype != 0:
    raise NotImplementedError("compressed (type %d) table block: tf.train.Saver never writes these" % ctype)
  return block


def _block_entries(block):
  n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
  end = len(block) - 4 - 4 * n_restarts
  pos, key = 0, b""
  while pos < end:
    shared, pos = _get_varint(block, pos)
    unshared, pos = _get_varint(block, pos)
    vlen, pos = _get_varint(block, pos)
    key = key[:shared] + bytes(block[pos:pos + unshared]); pos += unshared
    yield key, bytes(block[pos:pos + vlen])
    pos += vlen


def read_index(prefix):
  """-> (header dict, {name: entry dict}) of `<prefix>.index`."""
  with open(prefix + ".index", "rb") as f:
    data = f.read()
  if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != MAGIC:
    raise IOError("%s.index is not a TensorFlow checkpoint index (bad magic)" % prefix)
  footer = data[-48:]
  pos = 0
  _, pos = _get_varint(footer, pos); _, pos = _get_varint(footer, pos)      # metaindex handle
  ioff, pos = _get_varint(footer, pos); isize, pos = _get_varint(footer, pos)
  entries, header = {}, dict(num_shards=1, endianness=0)
  for _, handle in _block_entries(_read_block(data, ioff, isize)):
    boff, p = _get_varint(handle, 0)
    bsize, p = _get_varint(handle, p)
    for key, value in _block_entries(_read_block(data, boff, bsize)):
      if key == b"":
        for field, _, v in _parse_proto(value):
          if field == 1: header["num_shards"] = v
          elif field == 2: header["endianness"] = v
      else:
        entries[key.decode("utf-8")] = _parse_entry(value)
  if header["endianness"] != 0:
    raise NotImplementedError("big-endian checkpoint")
  return header, entries


def read_bundle(prefix, names=None):
  """All (or the named) tensors of the checkpoint `<prefix>` as {name: numpy array}."""
  header, entries = read_index(prefix)
  out, files = {}, {}
  try:
    for name, e in entries.items():
      if names is not None and name not in names:
        continue
      if e["slices"]:
        raise NotImplementedError("partitioned variable %s" % name)
      dt = DTYPES.get(e["dtype"])
      if dt is None:
        raise NotImplementedError("dtype enum %d of %s" % (e["dtype"], name))
      shard = e["shard_id"]
      if shard not in files:
        files[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, header["num_shards"]), "rb")
      f = files[shard]
      f.seek(e["offset"])
      raw = f.read(e["size"])
      want = int(np.prod(e["shape"], dtype=np.int64)) * np.dtype(dt).itemsize
      if len(raw) != e["size"] or e["size"] != want:
        raise IOError("checkpoint tensor %s: %d bytes on disk, shape %s needs %d" % (name, len(raw), e["shape"], want))
      out[name] = np.frombuffer(raw, dtype=np.dtype(dt).newbyteorder("<")).reshape(e["shape"]).astype(dt, copy=True)
  finally:
    for f in files.values():
      f.close()
  return out


def is_bundle(prefix):
  return os.path.exists(prefix + ".index")


# --------------------------------------------------------------------------- writer (same specification; tests)
def _shape_proto(shape):
  out = b""
  for d in shape:
    dim = b"\x08" + _put_varint(int(d) & ((1 << 64) - 1))
    out += b"\x12" + _put_varint(len(dim)) + dim
  return out


def _entry_proto(dtype, shape, offset, size, crc):
  shp = _shape_proto(shape)
  out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(shp)) + shp
  if offset:
    out += b"\x20" + _put_varint(offset)
  out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
  return out


def _build_block(items, restart_interval=16):
  buf, restarts, last = bytearray(), [], b""
  for i, (k, v) in enumerate(items):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(buf))
    else:
      while shared < min(len(k), len(last)) and k[shared] == last[shared]:
        shared += 1
    buf += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
    last = k
  if not restarts:
    restarts = [0]
  for r in restarts:
    buf += struct.pack("<I", r)
  buf += struct.pack("<I", len(restarts))
  return bytes(buf)


def write_bundle(prefix, tensors, block_entries=8):
  """Writes {name: array} as `<prefix>.index` + `<prefix>.data-00000-of-00001` (pure-Python crc32c: ~3 MB/s, meant
  for test-sized tensors)."""
  os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
  items = [(b"", b"\x08\x01\x10\x00\x1a\x02\x08\x01")]            # header: 1 shard, little endian, version {producer 1}
  offset = 0
  with open(prefix + ".data-00000-of-00001", "wb") as f:
    for name in sorted(tensors):
      a = np.asarray(tensors[name])
      a = a if a.flags.c_contiguous else a.copy()      # (ascontiguousarray would turn a scalar into shape (1,))
      raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
      f.write(raw)
      items.append((name.encode("utf-8"),
                    _entry_proto(DTYPE_ENUM[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
      offset += len(raw)
  out, index_items = bytearray(), []

  def emit(block):
    off = len(out)
    out.extend(block)
    out.append(0)
    out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
    return _put_varint(off) + _put_varint(len(block))

  for i in range(0, len(items), block_entries):
    chunk = items[i:i + block_entries]
    index_items.append((chunk[-1][0] + b"\x00", emit(_build_block(chunk))))
  meta = emit(_build_block([]))
  index = emit(_build_block(index_items, restart_interval=1))
  footer = meta + index
  footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
  out.extend(footer)
  with open(prefix + ".index", "wb") as f:
    f.write(bytes(out))
