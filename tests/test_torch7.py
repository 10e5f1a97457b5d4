"""Torch7 binary serialisation (catgen/torch7.py) -- CPU only.

There is no Torch7 and no .net fixture in this image: these tests pin the reader against byte strings assembled BY HAND from the
format's definition (File.lua / Tensor.c / Storage.c of torch7, restated in torch7.py's header), round-trip the writer through the
reader, and check that walking a saved module tree reproduces getParameters()'s order for the three architectures of the hot path.
"""
import struct

import numpy as np
import pytest

from catgen import torch7 as t7
from oracle import pyoracle as po


def i32(*v):
    return b"".join(struct.pack("<i", x) for x in v)


def i64(*v):
    return b"".join(struct.pack("<q", x) for x in v)


def s(x):
    return i32(len(x)) + x.encode()


def test_reader_on_hand_assembled_bytes():
    # torch.save of the table {epoch = 7, ok = true, name = "cat", w = FloatTensor(2,3) viewing elements 3.. of a 10-float storage}
    storage = np.arange(10, dtype=np.float32)
    tensor = (i32(4, 2) + s("V 1") + s("torch.FloatTensor") + i32(2) + i64(2, 3) + i64(3, 1) + i64(4)      # ndim, size, stride, offset (1-based)
              + i32(4, 3) + s("V 1") + s("torch.FloatStorage") + i64(10) + storage.tobytes())
    blob = (i32(3, 1, 4)
            + i32(2) + s("epoch") + i32(1) + struct.pack("<d", 7.0)
            + i32(2) + s("ok") + i32(5, 1)
            + i32(2) + s("name") + i32(2) + s("cat")
            + i32(2) + s("w") + tensor)
    r = t7.Reader(blob)
    tab = r.read()
    assert r.o == len(blob)
    assert tab["epoch"] == 7 and tab["ok"] is True and tab["name"] == "cat"
    assert np.array_equal(tab["w"].array(), storage[3:9].reshape(2, 3))
    # a second reference to the same tensor is an index only
    blob2 = i32(3, 1, 2) + i32(1) + struct.pack("<d", 1.0) + tensor + i32(1) + struct.pack("<d", 2.0) + i32(4, 2)
    tab2 = t7.Reader(blob2).read()
    assert tab2[1] is tab2[2]


def test_reader_handles_legacy_class_header_nil_storage_and_functions():
    empty = i32(4, 2) + s("torch.FloatTensor") + i32(0) + i64(1) + i32(0)          # no "V n" line; ndim 0, offset 1, nil storage
    fn = i32(8, 3) + i32(4) + b"\x1bLJ\x02" + i32(3, 4, 0)                           # a dumped function with an empty upvalue table
    blob = i32(3, 1, 2) + i32(2) + s("t") + empty + i32(2) + s("type") + fn
    tab = t7.Reader(blob).read()
    assert tab["t"].array().size == 0 and isinstance(tab["type"], t7.T7Function) and tab["type"].upvalues == {}


def test_reader_rejects_truncated_and_unknown():
    good = t7.dumps({"a": np.arange(6, dtype=np.float32)})
    with pytest.raises(t7.T7Error):
        t7.Reader(good[:-3]).read()
    with pytest.raises(t7.T7Error):
        t7.Reader(i32(42)).read()
    with pytest.raises(t7.T7Error):
        t7.Reader(i32(4, 1) + s("V 1") + s("torch.WeirdTensor")).read()
    bad_view = i32(4, 1) + s("V 1") + s("torch.FloatTensor") + i32(1) + i64(8) + i64(1) + i64(1) + i32(4, 2) + s("V 1") + s("torch.FloatStorage") + i64(4) + bytes(16)
    with pytest.raises(t7.T7Error):
        t7.Reader(bad_view).read().array()


def test_round_trip_scalars_tables_tensors(tmp_path):
    obj = {"epoch": 12, "lr": 1e-3, "flag": False, "none": None, "list": [1, 2, "x"], "nested": {"a": {"b": 2.5}},
           "f": np.random.default_rng(0).standard_normal((3, 4, 5)).astype(np.float32), "d": np.arange(5, dtype=np.float64),
           "l": np.arange(4, dtype=np.int64), "b": np.arange(7, dtype=np.uint8)}
    p = tmp_path / "x.t7"
    t7.save(str(p), obj)
    back = t7.load(str(p))
    assert back["epoch"] == 12 and back["lr"] == 1e-3 and back["flag"] is False and "none" in back and back["none"] is None
    assert back["list"] == {1: 1, 2: 2, 3: "x"} and back["nested"]["a"]["b"] == 2.5
    for k in ("f", "d", "l", "b"):
        a = back[k].array()
        assert a.dtype == obj[k].dtype and np.array_equal(a, obj[k])
    # strided views: a transposed view of a shared storage
    st = t7.T7Storage(np.arange(12, dtype=np.float32))
    v = t7.T7Tensor(st, (4, 3), (1, 4), 0)
    back = t7.Reader(t7.dumps({"v": v, "w": t7.T7Tensor(st, (2,), None, 10)})).read()
    assert np.array_equal(back["v"].array(), np.arange(12, dtype=np.float32).reshape(3, 4).T)
    assert np.array_equal(back["w"].array(), [10, 11]) and back["v"].storage is back["w"].storage


@pytest.mark.parametrize("kind,C", [(po.G32UPC, 3), (po.G32UP, 1)])
def test_G_tree_reproduces_getParameters_order(kind, C):
    m = po.Model(kind, C, 100, seed=3)
    flat = np.array(m.params)
    run = np.random.default_rng(1).uniform(0.5, 1.5, m.bn_running.size).astype(np.float32)
    tree = t7.tree_G(kind == po.G32UPC, C, 100, flat, run)
    back = t7.Reader(t7.dumps({"G": tree})).read()["G"]
    assert back.typename == "nn.Sequential"
    assert np.array_equal(t7.flat_parameters(back), flat)
    assert np.array_equal(t7.bn_running(back), run)
    ps = t7.parameters(back)
    assert len({id(p.storage) for p in ps}) == 1, "every weight is a view into ONE flat storage, as after getParameters()"
    assert ps[0].size == ((8192, 100) if kind == po.G32UPC else (8192, 100)) and ps[0].offset == 0 and ps[1].offset == 819200
    names = [k.typename for k in t7._children(back)]
    assert names[-2:] == ["cudnn.SpatialConvolution", "nn.Sigmoid"] and names.count("nn.SpatialUpSamplingNearest") == (3 if kind == po.G32UPC else 2)


def test_D_tree_reproduces_getParameters_order():
    m = po.Model(po.D32_ST3, 3, 100, seed=4)
    flat = np.array(m.params)
    back = t7.Reader(t7.dumps(t7.tree_D(3, flat))).read()
    assert np.array_equal(t7.flat_parameters(back), flat) and flat.size == 6664777
    assert t7.bn_running(back).size == 0
    kids = t7._children(back)
    assert kids[0].typename == "nn.Copy" and kids[-1].typename == "nn.Copy" and kids[8].typename == "nn.Concat" and kids[8]["dimension"] == 2
    assert [len(t7._children(b)) for b in t7._children(kids[8])] == [7, 7, 7, 6]
    # the localisation network's last Linear sits where models.lua:852-860 puts it: [nth, 64] weights then nth biases
    stn = kids[1]
    loc = t7._children(t7._children(t7._children(stn)[0])[1])[0]
    last = t7._children(loc)[-1]
    assert last.typename == "nn.Linear" and last["weight"].size == (1, 64) and last["bias"].size == (1,)
    with pytest.raises(t7.T7Error):
        t7.tree_D(3, flat[:-1])


def test_V_tree_and_legacy_running_std():
    flat, run = po.V32_init(3, seed=2)
    tree = t7.tree_V(3, flat, run)
    back = t7.Reader(t7.dumps(tree)).read()
    assert np.array_equal(t7.flat_parameters(back), flat) and np.array_equal(t7.bn_running(back), run)
    # nn before 2016: running_std = 1/sqrt(var + eps) instead of running_var
    for mod in t7._children(tree):
        if mod.typename.endswith("BatchNormalization"):
            var = mod.fields.pop("running_var").array().astype(np.float64)
            mod.fields["running_std"] = t7.T7Tensor.of((1.0 / np.sqrt(var + 1e-5)).astype(np.float32))
    old = t7.bn_running(t7.Reader(t7.dumps(tree)).read())
    assert np.abs(old - run).max() < 1e-5
