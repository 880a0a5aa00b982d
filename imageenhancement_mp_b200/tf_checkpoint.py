"""Pure-Python reader of TensorFlow object-graph checkpoints (``tf.train.Checkpoint`` / TensorBundle V2), so the
weights the reference saves at run_training.py:114-115,229 and restores at eval.py:112-118 can be evaluated here
without TensorFlow:  ``<prefix>.index`` + ``<prefix>.data-00000-of-0000N``  ->  ``{layer: (kernel HWIO, bias)}``.

STATUS - PARITY UNPINNED: TensorFlow is not installable in the build environment and the reference ships no
checkpoint, so this reader has never seen a TensorFlow-written file.  It restates the published on-disk formats
(below); ``tests/test_tf_checkpoint.py`` round-trips it against a writer restated from the same definitions, and pins
the pieces that TensorFlow-authored code in this image covers (tensorboard's copies of TensorShapeProto, the DataType
enum, TrackableObjectGraph and its crc32c / mask) against that code.  BundleEntryProto and the table layout remain
restated only.

Formats restated:
* ``.index`` is a LevelDB-format sorted string table (tensorflow/core/lib/io/table*, a fork of LevelDB's
  ``table/format.h`` / ``block.cc``): 48-byte footer = metaindex BlockHandle + index BlockHandle (varint64 offset,
  varint64 size each), zero padding to 40 bytes, magic ``0xdb4775248b80fb57`` little-endian.  Every block is followed
  by a 5-byte trailer (compression type, masked crc32c); TensorBundle writes with ``kNoCompression``
  (tensor_bundle.cc ``BundleWriter::Finish``).  A block is a run of prefix-compressed entries
  ``varint32 shared | varint32 non_shared | varint32 value_len | key_delta | value`` followed by the ``uint32`` restart
  offsets and their count.  The index block maps separator keys to the BlockHandles of the data blocks.
* Values are protobufs (tensor_bundle.proto): key ``""`` -> ``BundleHeaderProto{num_shards=1, endianness=2, version=3}``,
  every other key -> ``BundleEntryProto{dtype=1, shape=2 (TensorShapeProto{dim=2{size=1}}), shard_id=3, offset=4,
  size=5, crc32c=6 (fixed32), slices=7}``; tensor bytes sit raw (little-endian, row-major) at ``offset`` of data shard
  ``shard_id``.
* ``tf.train.Checkpoint(net=model)`` names a variable by the attribute path that reaches it:
  ``net/down1/conv2d1/kernel/.ATTRIBUTES/VARIABLE_VALUE``; optimizer slots carry ``.OPTIMIZER_SLOT`` and are skipped.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5
VARIABLE_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow/core/framework/types.proto
DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
          6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("bool"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
          22: np.dtype("<u4"), 23: np.dtype("<u8")}


class CheckpointFormatError(ValueError):
    pass


# ------------------------------------------------------------------ varints / protobuf wire format
def read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointFormatError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointFormatError("varint longer than 64 bits")


def parse_proto(buf):
    """Minimal protobuf decoder: {field number: [values]}; varint -> int, fixed32/64 -> int, length-delimited -> bytes."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = read_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = read_varint(buf, pos)
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            n, pos = read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            if len(v) != n:
                raise CheckpointFormatError("truncated length-delimited field")
            pos += n
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointFormatError(f"unsupported protobuf wire type {wire}")
        out.setdefault(field, []).append(v)
    return out


# ------------------------------------------------------------------ crc32c (Castagnoli), masked as LevelDB / TensorFlow store it
_CRC_TABLE = None
CRC_MASK_DELTA = 0xA282EAD8


def crc32c(data):
    """CRC-32C of a bytes-like object (tensorflow/core/lib/hash/crc32c.h); table-driven, for index blocks and small
    tensors - pure Python, about 5 MB/s."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        table = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            table.append(c)
        _CRC_TABLE = table
    table = _CRC_TABLE
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = table[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    """crc32c::Mask: rotate right by 15 bits and add a constant, so that a CRC of data that embeds CRCs stays useful."""
    return (((crc >> 15) | (crc << 17)) + CRC_MASK_DELTA) & 0xFFFFFFFF


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_shape(buf):
    """TensorShapeProto -> tuple of ints."""
    msg = parse_proto(buf)
    if msg.get(3, [0])[0]:
        raise CheckpointFormatError("tensor of unknown rank in a checkpoint")
    dims = []
    for d in msg.get(2, []):
        dims.append(_signed64(parse_proto(d).get(1, [0])[0]))
    return tuple(dims)


# ------------------------------------------------------------------ LevelDB-format table
def _read_block(buf, offset, size, verify=True):
    end = offset + size
    if end + BLOCK_TRAILER_LEN > len(buf):
        raise CheckpointFormatError("block handle points past the end of the index file")
    ctype = buf[end]
    if ctype == 1:
        raise CheckpointFormatError("snappy-compressed table block: TensorBundle indexes are written uncompressed; "
                                    "this file was produced by something else")
    if ctype != 0:
        raise CheckpointFormatError(f"unknown block compression type {ctype}")
    if verify:         # trailer: masked crc32c over the block contents and the compression-type byte (table/format.cc)
        stored = struct.unpack_from("<I", buf, end + 1)[0]
        if stored != mask_crc(crc32c(buf[offset:end + 1])):
            raise CheckpointFormatError(f"checksum mismatch in the table block at offset {offset}: the index file is corrupt")
    return buf[offset:end]


def _block_entries(block):
    """(key, value) pairs of one table block, undoing the key prefix compression."""
    if len(block) < 4:
        raise CheckpointFormatError("table block shorter than its restart count")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointFormatError("corrupt restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = read_varint(block, pos)
        non_shared, pos = read_varint(block, pos)
        vlen, pos = read_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointFormatError("corrupt table entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order.  ``verify``: check every block's masked
    crc32c trailer."""
    with open(path, "rb") as fh:
        buf = fh.read()
    if len(buf) < FOOTER_LEN:
        raise CheckpointFormatError(f"{path}: shorter than a table footer")
    footer = buf[-FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, FOOTER_LEN - 8)[0] != TABLE_MAGIC:
        raise CheckpointFormatError(f"{path}: not a TensorFlow / LevelDB table (bad magic)")
    pos = 0
    _, pos = read_varint(footer, pos)            # metaindex handle (unused)
    _, pos = read_varint(footer, pos)
    idx_off, pos = read_varint(footer, pos)
    idx_size, pos = read_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, idx_off, idx_size, verify)):
        off, p = read_varint(handle, 0)
        size, _ = read_varint(handle, p)
        out.extend(_block_entries(_read_block(buf, off, size, verify)))
    return out


# ------------------------------------------------------------------ TensorBundle
def read_bundle(prefix, keys=None, verify="index"):
    """``{key: numpy array}`` of the numeric tensors of the bundle ``prefix`` (``keys``: optional filter callable).
    String / variant tensors (the object graph itself) are skipped.

    ``verify``: "index" (default) checks the masked crc32c trailer of every index block; "all" also checks every
    tensor's bytes against ``BundleEntryProto.crc32c`` (pure Python, ~5 MB/s: minutes for the reference's 200 MB of
    weights - meant for a one-off integrity check); "none" checks nothing."""
    if verify not in ("index", "all", "none"):
        raise ValueError("verify must be 'index', 'all' or 'none'")
    entries = read_table(prefix + ".index", verify=verify != "none")
    if not entries or entries[0][0] != b"":
        raise CheckpointFormatError("bundle index has no header entry")
    header = parse_proto(entries[0][1])
    num_shards = header.get(1, [1])[0]
    if header.get(2, [0])[0] != 0:
        raise CheckpointFormatError("big-endian bundle")
    shards = {}
    out = {}
    for raw_key, value in entries[1:]:
        key = raw_key.decode("utf-8")
        if keys is not None and not keys(key):
            continue
        e = parse_proto(value)
        if 7 in e:
            raise CheckpointFormatError(f"{key}: sliced (partitioned) variables are not supported")
        dtype = DTYPES.get(e.get(1, [0])[0])
        if dtype is None:
            continue                                           # DT_STRING, DT_VARIANT, ... : not weights
        shape = parse_shape(e[2][0]) if 2 in e else ()
        shard = e.get(3, [0])[0]
        off, size = e.get(4, [0])[0], e.get(5, [0])[0]
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if size != count * dtype.itemsize:
            raise CheckpointFormatError(f"{key}: {size} bytes on disk for shape {shape} of {dtype}")
        if shard not in shards:
            path = "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)
            if not os.path.exists(path):
                raise CheckpointFormatError(f"data shard {path} is missing")
            shards[shard] = np.memmap(path, dtype=np.uint8, mode="r")
        data = shards[shard]
        if off + size > data.shape[0]:
            raise CheckpointFormatError(f"{key}: tensor bytes run past the end of the data shard")
        raw = data[off:off + size].tobytes()
        if verify == "all" and 6 in e and e[6][0] != mask_crc(crc32c(raw)):
            raise CheckpointFormatError(f"{key}: checksum mismatch - the tensor bytes in the data shard are corrupt")
        out[key] = np.frombuffer(raw, dtype=dtype).reshape(shape).copy()
    return out


def latest_checkpoint(directory):
    """``tf.train.latest_checkpoint``: the prefix named by the ``checkpoint`` state file (text CheckpointState proto:
    ``model_checkpoint_path: "ckpt-12"``), or None."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as fh:
        for line in fh:
            if line.startswith("model_checkpoint_path:"):
                name = line.split(":", 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(directory, name)
    return None


def _resolve_path(path, layer_order):
    """Attribute path of a variable.  Keras also reaches every layer through ``layer_with_weights-N`` edges (creation
    order); a checkpoint saver that happened to name variables through them is mapped back: the top-level index into
    ``layer_order`` (the model's weight-bearing attributes in creation order), a nested one into ``conv2d<N+1>`` (the
    blocks of model_library.py:65-101 create conv2d1..3 in that order)."""
    out = []
    for depth, comp in enumerate(path):
        if comp.startswith("layer_with_weights-"):
            idx = int(comp.rsplit("-", 1)[1])
            if depth == 0:
                if layer_order is None or idx >= len(layer_order):
                    raise CheckpointFormatError(f"variable path {'/'.join(path)} needs the model's layer order")
                comp = layer_order[idx]
            else:
                comp = "conv2d%d" % (idx + 1)
        out.append(comp)
    return out


def load_tf_checkpoint(prefix, root="net", dtype=np.float32, layer_order=None):
    """``{layer name: (kernel [kh,kw,cin,cout], bias [cout])}`` (torch tensors, the format of ``weights.load_npz``) from
    the variables under ``root`` - the keyword the model was given in ``tf.train.Checkpoint`` (``net`` at
    run_training.py:114; eval.py:112 uses the same object).  Layer names are the Keras attribute paths joined with
    ``.`` (``down1.conv2d1``, ``layer0``, ...), as everywhere in this package.  ``layer_order``: see _resolve_path."""
    import torch
    want = root + "/"
    tensors = read_bundle(prefix, keys=lambda k: k.startswith(want) and k.endswith(VARIABLE_SUFFIX)
                          and ".OPTIMIZER_SLOT" not in k)
    kernels, biases = {}, {}
    for key, arr in tensors.items():
        path = key[len(want):-len(VARIABLE_SUFFIX)].split("/")
        if len(path) < 2 or path[-1] not in ("kernel", "bias"):
            continue
        path = _resolve_path(path[:-1], layer_order) + path[-1:]
        (kernels if path[-1] == "kernel" else biases)[".".join(path[:-1])] = arr
    if not kernels:
        raise CheckpointFormatError(
            f"no '{root}/<layer>/kernel{VARIABLE_SUFFIX}' variables in {prefix}.index (is the model stored under "
            f"another keyword than '{root}'?)")
    out = {}
    for name, k in kernels.items():
        if k.ndim != 4:
            raise CheckpointFormatError(f"{name}: kernel of rank {k.ndim}, expected HWIO")
        b = biases.get(name)
        if b is None:
            b = np.zeros(k.shape[-1], dtype=k.dtype)
        if b.shape != (k.shape[-1],):
            raise CheckpointFormatError(f"{name}: bias shape {b.shape} does not match {k.shape[-1]} output channels")
        out[name] = (torch.from_numpy(np.ascontiguousarray(k, dtype=dtype)), torch.from_numpy(np.ascontiguousarray(b, dtype=dtype)))
    return out
