"""Reader for the reference's pretrained policy without TensorFlow.

The reference saves its four MADDPG networks with ``tf.train.Saver`` (main.py:209, 541-543); the
shipped bundle is ``model_data/baseline/66.cptk.{index,data-00000-of-00001}``.  The format:

* ``.index`` is a LevelDB-style sorted table: data blocks of prefix-compressed (key, value) entries
  followed by a restart array, an index block, and a 48-byte footer (two block handles + the magic
  ``0xdb4775248b80fb57``).  A block is followed by a 1-byte compression type and a 4-byte CRC.
* every value is a ``BundleEntryProto``: 1 dtype, 2 shape (``TensorShapeProto``: repeated field 2
  ``Dim`` with field 1 ``size``), 3 shard, 4 offset, 5 size, 6 crc32c.
* ``.data-00000-of-00001`` holds the raw little-endian tensors at (offset, size).

Only what the actor needs is implemented: uncompressed blocks, float32 tensors, a single shard.
"""
import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DT_FLOAT = 1


def _varint(buf, pos):
    out = shift = 0
    while True:
        byte = buf[pos]
        pos += 1
        out |= (byte & 0x7F) << shift
        if not byte & 0x80:
            return out, pos
        shift += 7


def _block(buf, offset, size):
    """Entries of the table block at (offset, size)."""
    if buf[offset + size] != 0:
        raise ValueError("compressed table block (type %d) is not supported" % buf[offset + size])
    raw = buf[offset:offset + size]
    n_restart = struct.unpack("<I", raw[-4:])[0]
    end = len(raw) - 4 - 4 * n_restart
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(raw, pos)
        unshared, pos = _varint(raw, pos)
        vlen, pos = _varint(raw, pos)
        key = key[:shared] + raw[pos:pos + unshared]
        pos += unshared
        out.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return out


def _fields(buf):
    """(field number, wire type, value) of a protobuf message; value is an int or bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        num, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
        elif wire == 2:
            n, pos = _varint(buf, pos)
            val = buf[pos:pos + n]
            pos += n
        elif wire == 5:
            val = struct.unpack("<I", buf[pos:pos + 4])[0]
            pos += 4
        elif wire == 1:
            val = struct.unpack("<Q", buf[pos:pos + 8])[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        yield num, wire, val


def read_bundle_index(index_path):
    """name -> dict(dtype, shape, shard, offset, size) for every tensor of the bundle."""
    buf = open(index_path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _MAGIC:
        raise ValueError("%s is not a tensor bundle index" % index_path)
    foot = buf[-48:]
    _, pos = _varint(foot, 0)           # metaindex handle
    _, pos = _varint(foot, pos)
    idx_off, pos = _varint(foot, pos)
    idx_size, pos = _varint(foot, pos)
    out = {}
    for _, handle in _block(buf, idx_off, idx_size):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, val in _block(buf, off, size):
            if not key:                 # header entry (BundleHeaderProto)
                continue
            ent = {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0}
            for num, _, v in _fields(val):
                if num == 1:
                    ent["dtype"] = v
                elif num == 2:
                    dims = []
                    for n2, _, v2 in _fields(v):
                        if n2 == 2:
                            dims.append(next((v3 for n3, _, v3 in _fields(v2) if n3 == 1), 0))
                    ent["shape"] = tuple(dims)
                elif num == 3:
                    ent["shard"] = v
                elif num == 4:
                    ent["offset"] = v
                elif num == 5:
                    ent["size"] = v
            out[key.decode()] = ent
    return out


def read_bundle(prefix, names=None):
    """name -> float32 array for the tensors ``names`` (default: all float32 tensors) of the bundle
    ``prefix`` (``prefix.index`` + ``prefix.data-00000-of-00001``)."""
    index = read_bundle_index(prefix + ".index")
    data_path = prefix + ".data-00000-of-00001"
    out = {}
    with open(data_path, "rb") as f:
        for name, ent in index.items():
            if names is not None and name not in names:
                continue
            if ent["dtype"] != _DT_FLOAT or ent["shard"] != 0:
                if names is not None:
                    raise ValueError("tensor %s: dtype %d / shard %d not supported" % (name, ent["dtype"], ent["shard"]))
                continue
            f.seek(ent["offset"])
            raw = f.read(ent["size"])
            arr = np.frombuffer(raw, dtype="<f4").astype(np.float32)
            out[name] = arr.reshape(ent["shape"])
    if names is not None:
        missing = [n for n in names if n not in out]
        if missing:
            raise KeyError("not in %s: %s" % (prefix, ", ".join(missing)))
    return out


def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint (main.py:541): the prefix named by the ``checkpoint`` file."""
    with open(os.path.join(model_dir, "checkpoint")) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                name = line.split(":", 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(model_dir, name)
    raise FileNotFoundError("no model_checkpoint_path in %s/checkpoint" % model_dir)
