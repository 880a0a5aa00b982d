"""tf_checkpoint.py (reader of TensorFlow object-graph checkpoints, eval.py:112-118 / run_training.py:114,229) against a
WRITER restated here from the same published format definitions (LevelDB-format table + tensor_bundle.proto).
PARITY UNPINNED: no TensorFlow in this environment and no checkpoint in the reference, so neither side has been checked
against a TensorFlow-written file; the test pins the reader's handling of prefix compression, multi-block indexes,
shards, dtypes, skipped entries and the reference's variable naming."""
import os
import struct

import numpy as np
import pytest
import torch

from imageenhancement_mp_b200 import tf_checkpoint as tfc


# ------------------------------------------------------------------ a minimal TensorBundle writer
def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def field(num, wire, payload):
    return varint((num << 3) | wire) + payload


def crc32c(data):
    table = getattr(crc32c, "t", None)
    if table is None:
        table = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            table.append(c)
        crc32c.t = table
    c = 0xFFFFFFFF
    for b in data:
        c = table[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked(crc):
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def build_block(entries, restart_interval):
    buf, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        buf += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def write_table(path, entries, block_size=256):
    out = bytearray()
    handles = []

    def emit(block):
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", masked(crc32c(block + b"\x00"))))
        return varint(off) + varint(len(block))

    cur, cur_size = [], 0
    for k, v in entries:
        cur.append((k, v))
        cur_size += len(k) + len(v)
        if cur_size >= block_size:
            handles.append((cur[-1][0], emit(build_block(cur, 16))))
            cur, cur_size = [], 0
    if cur:
        handles.append((cur[-1][0], emit(build_block(cur, 16))))
    meta = emit(build_block([], 1))
    index = emit(build_block(handles, 1))
    footer = meta + index
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", tfc.TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as fh:
        fh.write(out)


DT = {np.dtype("<f4"): 1, np.dtype("<f8"): 2, np.dtype("<i4"): 3, np.dtype("<i8"): 9}


def write_bundle(prefix, tensors, strings=(), num_shards=1, shard_of=lambda key: 0, block_size=256):
    """tensors: {key: numpy array}; strings: keys stored as DT_STRING entries (skipped by the reader)."""
    data = [bytearray() for _ in range(num_shards)]
    entries = [(b"", field(1, 0, varint(num_shards)) + field(3, 2, varint(2) + field(1, 0, varint(1))))]
    for key in sorted(list(tensors) + list(strings), key=lambda s: s.encode()):
        if key in strings:
            shard = 0
            payload = b"\x05hello"
            e = field(1, 0, varint(7)) + field(2, 2, varint(0)) + field(4, 0, varint(len(data[0]))) + \
                field(5, 0, varint(len(payload))) + field(6, 5, struct.pack("<I", masked(crc32c(payload))))
            data[0] += payload
        else:
            arr = tensors[key] if tensors[key].ndim == 0 else np.ascontiguousarray(tensors[key])   # (0-d stays 0-d)
            shard = shard_of(key)
            shape = b"".join(field(2, 2, (lambda d: varint(len(d)) + d)(field(1, 0, varint(s)))) for s in arr.shape)
            raw = arr.tobytes()
            e = field(1, 0, varint(DT[arr.dtype])) + field(2, 2, varint(len(shape)) + shape)
            if shard:
                e += field(3, 0, varint(shard))
            e += field(4, 0, varint(len(data[shard]))) + field(5, 0, varint(len(raw))) + \
                field(6, 5, struct.pack("<I", masked(crc32c(raw))))
            data[shard] += raw
        entries.append((key.encode(), e))
    write_table(prefix + ".index", entries, block_size)
    for i, d in enumerate(data):
        with open("%s.data-%05d-of-%05d" % (prefix, i, num_shards), "wb") as fh:
            fh.write(d)


SUF = tfc.VARIABLE_SUFFIX


def reference_style_tensors(rng):
    layers = {"layer0": (3, 5, 8), "down1/conv2d1": (3, 8, 8), "down1/conv2d2": (3, 8, 8), "Coef_up1/conv2d3": (3, 16, 4),
              "layer3_1": (2, 4, 6)}
    t = {}
    for name, (k, cin, cout) in layers.items():
        t[f"net/{name}/kernel{SUF}"] = rng.standard_normal((k, k, cin, cout)).astype("<f4")
        t[f"net/{name}/bias{SUF}"] = rng.standard_normal(cout).astype("<f4")
        t[f"net/{name}/kernel/.OPTIMIZER_SLOT/optimizer/m{SUF}"] = np.zeros((k, k, cin, cout), "<f4")
        t[f"net/{name}/kernel/.OPTIMIZER_SLOT/optimizer/v{SUF}"] = np.ones((k, k, cin, cout), "<f4")
    t[f"step{SUF}"] = np.array(7, "<i4")
    t[f"iterate{SUF}"] = np.array(12345678901, "<i8")
    t[f"optimizer/beta_1{SUF}"] = np.array(0.9, "<f4")
    t[f"optimizer/iter{SUF}"] = np.array(4000, "<i8")
    return layers, t


@pytest.mark.parametrize("num_shards,block_size", [(1, 64), (1, 4096), (2, 256)])
def test_bundle_roundtrip_reference_naming(tmp_path, num_shards, block_size):
    rng = np.random.default_rng(3)
    layers, t = reference_style_tensors(rng)
    prefix = str(tmp_path / "ckpt-3")
    write_bundle(prefix, t, strings=("_CHECKPOINTABLE_OBJECT_GRAPH",), num_shards=num_shards,
                 shard_of=lambda key: hash(key) % num_shards if num_shards > 1 else 0, block_size=block_size)
    got = tfc.read_bundle(prefix)
    assert set(got) == set(t)                                  # the DT_STRING object graph is skipped
    for k, v in t.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    W = tfc.load_tf_checkpoint(prefix)                         # root "net", run_training.py:114
    assert sorted(W) == sorted(n.replace("/", ".") for n in layers)
    for name in layers:
        k, b = W[name.replace("/", ".")]
        assert isinstance(k, torch.Tensor) and k.dtype == torch.float32
        assert np.array_equal(k.numpy(), t[f"net/{name}/kernel{SUF}"]) and np.array_equal(b.numpy(), t[f"net/{name}/bias{SUF}"])
    with pytest.raises(tfc.CheckpointFormatError, match="another keyword"):
        tfc.load_tf_checkpoint(prefix, root="myAwesomeModel")


def test_weights_dict_is_the_npz_format(tmp_path):
    """A checkpoint of (a slice of) the real layer list loads into the same dict weights.load_npz gives."""
    from imageenhancement_mp_b200 import synth, weights
    params = dict(synth.DEFAULT_PARAMS)
    layers = [l for l in weights.simplemodel_layers(params) if l[2] * l[3] <= 64 * 128][:6]
    W = weights.init_weights(layers)
    t = {}
    for name, (k, b) in W.items():
        t["net/" + name.replace(".", "/") + "/kernel" + SUF] = k.numpy()
        t["net/" + name.replace(".", "/") + "/bias" + SUF] = b.numpy()
    prefix = str(tmp_path / "ckpt-1")
    write_bundle(prefix, t, block_size=512)
    got = tfc.load_tf_checkpoint(prefix)
    npz = str(tmp_path / "w.npz")
    weights.save_npz(npz, W)
    ref = weights.load_npz(npz)
    assert sorted(got) == sorted(ref)
    for name in ref:
        assert torch.equal(got[name][0], ref[name][0]) and torch.equal(got[name][1], ref[name][1])
    # weights.load() dispatch: prefix, .index file, directory with a CheckpointManager state file, .npz
    with open(str(tmp_path / "checkpoint"), "w") as fh:
        fh.write('model_checkpoint_path: "ckpt-1"\n')
    for path in (prefix, prefix + ".index", str(tmp_path), npz):
        again = weights.load(path)
        assert sorted(again) == sorted(ref) and all(torch.equal(again[n][0], ref[n][0]) for n in ref)
    with pytest.raises(FileNotFoundError):
        weights.load(str(tmp_path / "nothing-here"))


def test_layer_with_weights_aliases(tmp_path):
    """Variables named through Keras' creation-order edges map back to the attribute names."""
    from imageenhancement_mp_b200 import synth, weights
    layers = weights.simplemodel_layers(dict(synth.DEFAULT_PARAMS))
    rng = np.random.default_rng(8)
    t = {f"net/layer_with_weights-0/kernel{SUF}": rng.standard_normal((3, 3, 5, 4)).astype("<f4"),
         f"net/layer_with_weights-0/bias{SUF}": rng.standard_normal(4).astype("<f4"),
         f"net/layer_with_weights-1/layer_with_weights-1/kernel{SUF}": rng.standard_normal((3, 3, 4, 4)).astype("<f4"),
         f"net/layer_with_weights-5/conv2d3/kernel{SUF}": rng.standard_normal((3, 3, 4, 2)).astype("<f4")}
    prefix = str(tmp_path / "ckpt-2")
    write_bundle(prefix, t)
    W = weights.load(prefix, layers=layers)
    assert sorted(W) == ["Coef_up1.conv2d3", "down1.conv2d2", "layer0"]
    assert np.array_equal(W["down1.conv2d2"][0].numpy(), t[f"net/layer_with_weights-1/layer_with_weights-1/kernel{SUF}"])
    assert float(W["down1.conv2d2"][1].abs().sum()) == 0.0            # no bias stored: zeros
    with pytest.raises(tfc.CheckpointFormatError, match="layer order"):
        tfc.load_tf_checkpoint(prefix)


def test_latest_checkpoint_and_errors(tmp_path):
    d = str(tmp_path)
    assert tfc.latest_checkpoint(d) is None
    with open(os.path.join(d, "checkpoint"), "w") as fh:
        fh.write('model_checkpoint_path: "ckpt-12"\nall_model_checkpoint_paths: "ckpt-11"\nall_model_checkpoint_paths: "ckpt-12"\n')
    assert tfc.latest_checkpoint(d) == os.path.join(d, "ckpt-12")
    bad = os.path.join(d, "bad")
    with open(bad + ".index", "wb") as fh:
        fh.write(b"\x00" * 64)
    with pytest.raises(tfc.CheckpointFormatError, match="bad magic"):
        tfc.read_bundle(bad)
    rng = np.random.default_rng(1)
    _, t = reference_style_tensors(rng)
    prefix = os.path.join(d, "ckpt-12")
    write_bundle(prefix, t)
    os.remove(prefix + ".data-00000-of-00001")
    with pytest.raises(tfc.CheckpointFormatError, match="missing"):
        tfc.read_bundle(prefix)


# ------------------------------------------------------------------ pins against TensorFlow-authored code in this image
# tensorboard ships TensorFlow's own protobuf definitions (tensorboard.compat.proto) and its own crc32c / mask
# (tensorboard.compat.tensorflow_stub.pywrap_tensorflow, used for event files): the parts of the checkpoint format
# they cover are checked against them.  BundleEntryProto / the table format itself are not in tensorboard.
def test_crc32c_and_mask_match_tensorflows_own():
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    assert tfc.crc32c(b"123456789") == 0xE3069283                      # the CRC-32C check value (RFC 3720 B.4)
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 64, 1000):
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert tfc.crc32c(data) == pw.crc32c(data) == crc32c(data)
        assert tfc.mask_crc(tfc.crc32c(data)) == pw.masked_crc32c(data) == masked(crc32c(data))


def test_shape_and_dtype_decoding_match_tensorflows_protos():
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    for dims in [(), (64,), (3, 3, 5, 64), (3, 3, 2048, 512), (0, 7), (1 << 40, 2)]:
        msg = shape_pb2.TensorShapeProto(dim=[shape_pb2.TensorShapeProto.Dim(size=d) for d in dims])
        assert tfc.parse_shape(msg.SerializeToString()) == dims
    with pytest.raises(tfc.CheckpointFormatError):
        tfc.parse_shape(shape_pb2.TensorShapeProto(unknown_rank=True).SerializeToString())
    names = {1: "DT_FLOAT", 2: "DT_DOUBLE", 3: "DT_INT32", 4: "DT_UINT8", 5: "DT_INT16", 6: "DT_INT8", 9: "DT_INT64",
             10: "DT_BOOL", 17: "DT_UINT16", 19: "DT_HALF", 22: "DT_UINT32", 23: "DT_UINT64"}
    assert set(names) == set(tfc.DTYPES)
    np_of = {"DT_FLOAT": "<f4", "DT_DOUBLE": "<f8", "DT_INT32": "<i4", "DT_UINT8": "u1", "DT_INT16": "<i2", "DT_INT8": "i1",
             "DT_INT64": "<i8", "DT_BOOL": "bool", "DT_UINT16": "<u2", "DT_HALF": "<f2", "DT_UINT32": "<u4",
             "DT_UINT64": "<u8"}
    for num, name in names.items():
        assert types_pb2.DataType.Value(name) == num and tfc.DTYPES[num] == np.dtype(np_of[name])
    assert types_pb2.DataType.Value("DT_STRING") not in tfc.DTYPES            # the object graph entry is skipped


def test_object_graph_entry_is_skipped_and_its_keys_are_the_ones_read(tmp_path):
    """A tf.train.Checkpoint stores its TrackableObjectGraph (proto from tensorboard.compat.proto) as a DT_STRING entry
    `_CHECKPOINTABLE_OBJECT_GRAPH`; each variable's SerializedTensor.checkpoint_key is the key of its bundle entry.
    Build that graph for net.down1.conv2d1.{kernel,bias}, store it next to the tensors, and read the bundle back."""
    tog = pytest.importorskip("tensorboard.compat.proto.trackable_object_graph_pb2")
    g = tog.TrackableObjectGraph()
    root, net, down1, conv, kernel, bias = (g.nodes.add() for _ in range(6))
    root.children.add(node_id=1, local_name="net")
    net.children.add(node_id=2, local_name="down1")
    down1.children.add(node_id=3, local_name="conv2d1")
    conv.children.add(node_id=4, local_name="kernel")
    conv.children.add(node_id=5, local_name="bias")
    keys = {}
    for node, leaf in ((kernel, "kernel"), (bias, "bias")):
        key = "net/down1/conv2d1/%s/.ATTRIBUTES/VARIABLE_VALUE" % leaf
        node.attributes.add(name="VARIABLE_VALUE", full_name="down1/conv2d1/" + leaf, checkpoint_key=key)
        keys[leaf] = key
        assert key.endswith(tfc.VARIABLE_SUFFIX)
    rng = np.random.default_rng(2)
    tensors = {keys["kernel"]: rng.standard_normal((3, 3, 64, 64)).astype("<f4"),
               keys["bias"]: rng.standard_normal(64).astype("<f4")}
    prefix = str(tmp_path / "ckpt-3")
    write_bundle(prefix, tensors, strings=("_CHECKPOINTABLE_OBJECT_GRAPH",))
    got = tfc.read_bundle(prefix, verify="all")
    assert sorted(got) == sorted(tensors)
    for k in tensors:
        assert np.array_equal(got[k], tensors[k])
    W = tfc.load_tf_checkpoint(prefix)
    assert list(W) == ["down1.conv2d1"] and tuple(W["down1.conv2d1"][0].shape) == (3, 3, 64, 64)


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(4)
    t = {"net/layer0/kernel" + SUF: rng.standard_normal((3, 3, 5, 64)).astype("<f4"),
         "net/layer0/bias" + SUF: rng.standard_normal(64).astype("<f4")}
    prefix = str(tmp_path / "ckpt-1")
    write_bundle(prefix, t)
    assert len(tfc.read_bundle(prefix, verify="all")) == 2
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[100] ^= 0x10
    open(data, "wb").write(raw)
    assert len(tfc.read_bundle(prefix)) == 2                        # default: the index only
    with pytest.raises(tfc.CheckpointFormatError, match="checksum"):
        tfc.read_bundle(prefix, verify="all")
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 0x01
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(tfc.CheckpointFormatError):
        tfc.read_bundle(prefix)
