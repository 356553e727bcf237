"""Torch7 binary format (bot7_b200/t7.py): result persistence of bot:save (bots/abstract.lua:234-240) and the data
files of examples/ -- SURVEY.md section 8(f) row 4.  Pinned against the reference's own fixture."""
import hashlib
import os
import struct

import numpy as np
import pytest

from bot7_b200 import t7

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "ref_iris_test30.npz")      # tests/golden/make_t7_fixture.py


def reference_table():
    z = np.load(FIXTURE)
    return {str(k): z[str(k)] for k in z["order"]}, str(z["sha256"]), int(z["nbytes"])


def test_writer_reproduces_the_reference_file_byte_for_byte():
    # the arrays were read from the reference's examples/data/iris_test30.t7 by this reader; writing them back must give
    # the very bytes torch.save wrote (compared through their SHA-256 and length: the file itself is not vendored)
    table, sha, nbytes = reference_table()
    raw = t7.dumps(table)
    assert len(raw) == nbytes == 6528 and hashlib.sha256(raw).hexdigest() == sha
    assert sha == "464c79fed7468d631f45818436736bc3b646cdd7ac9d3a99ac8bca6ae246264a"


def test_reads_the_reference_file_content():
    table, _, _ = reference_table()
    o = t7.loads(t7.dumps(table))                                      # == the reference bytes (previous test)
    assert list(o) == ["ye", "xe", "yr", "xr"]                       # file order of the Lua table
    assert o["xe"].shape == (30, 4) and o["xr"].shape == (120, 4) and o["ye"].shape == (30,) and o["yr"].shape == (120,)
    assert all(v.dtype == np.float64 for v in o.values())
    assert np.array_equal(o["xe"][0], [5.5, 4.2, 1.4, 0.2]) and np.array_equal(o["ye"][:6], [1, 2, 3, 2, 3, 3])
    assert set(np.unique(np.concatenate([o["ye"], o["yr"]]))) == {1.0, 2.0, 3.0}      # iris classes, 1-based
    assert np.bincount(np.concatenate([o["ye"], o["yr"]]).astype(int)).tolist() == [0, 50, 50, 50]
    for k in o:
        assert np.array_equal(o[k], table[k])


def test_round_trip_of_every_supported_value():
    r = np.random.default_rng(0)
    shared = {"a": 1.5}
    obj = {"best": {"x": r.random((1, 6)), "y": r.random((1, 1)), "t": 17}, "x": r.random((20, 6)), "y": r.random((20, 1)),
           "name": "bayesopt", "flag": True, "off": False, "nothing": None, 3: "three", 2.5: -1.0,
           "list": [1.0, "two", {"k": np.arange(5, dtype=np.int64)}],
           "f32": r.random((3, 2, 2)).astype(np.float32), "i32": np.arange(-3, 3, dtype=np.int32), "u8": np.arange(7, dtype=np.uint8),
           "empty": np.empty((0,)), "s1": shared, "s2": shared}
    back = t7.loads(t7.dumps(obj))
    assert back["s1"] is back["s2"] and back["s1"] == {"a": 1.5}                      # shared reference survives
    assert back["nothing"] is None and back["flag"] is True and back["off"] is False and back[3] == "three" and back[2.5] == -1.0
    assert back["best"]["t"] == 17.0 and back["name"] == "bayesopt"
    assert back["list"][1] == 1.0 and back["list"][2] == "two" and np.array_equal(back["list"][3]["k"], np.arange(5))
    for k in ("x", "y", "f32", "i32", "u8"):
        assert back[k].dtype == obj[k].dtype and np.array_equal(back[k], obj[k])
    assert np.array_equal(back["best"]["x"], obj["best"]["x"]) and back["empty"].size == 0
    assert t7.dumps(back) == t7.dumps(obj)                                            # idempotent


def test_strided_views_and_offsets_are_honoured():
    # a 3 x 2 tensor viewing a 10-element storage with offset 2 (1-based 3) and strides (3, 1), written by hand
    st = np.arange(10, dtype=np.float64)
    b = struct.pack("<ii", 4, 1) + struct.pack("<i", 3) + b"V 1" + struct.pack("<i", 18) + b"torch.DoubleTensor"
    b += struct.pack("<i", 2) + struct.pack("<qq", 3, 2) + struct.pack("<qq", 3, 1) + struct.pack("<q", 3)
    b += struct.pack("<ii", 4, 2) + struct.pack("<i", 3) + b"V 1" + struct.pack("<i", 19) + b"torch.DoubleStorage"
    b += struct.pack("<q", 10) + st.tobytes()
    a = t7.loads(b)
    assert np.array_equal(a, [[2, 3], [5, 6], [8, 9]])
    # a transposed array is written contiguously and reads back equal
    m = np.arange(12, dtype=np.float64).reshape(3, 4).T
    assert np.array_equal(t7.loads(t7.dumps(m)), m)


def test_malformed_streams_fail_loudly():
    good = t7.dumps({"x": np.arange(4.0)})
    with pytest.raises(t7.T7Error, match="truncated"):
        t7.loads(good[:-3])
    with pytest.raises(t7.T7Error, match="trailing"):
        t7.loads(good + b"\0")
    with pytest.raises(t7.T7Error, match="type tag"):
        t7.loads(struct.pack("<i", 6))                                # a serialised Lua function is not data
    with pytest.raises(t7.T7Error, match="exceeds"):
        bad = bytearray(good)
        i = bad.index(b"torch.DoubleTensor") + 18 + 4                 # first size field
        bad[i:i + 8] = struct.pack("<q", 400)
        t7.loads(bytes(bad))
    with pytest.raises(t7.T7Error, match="cannot serialise"):
        t7.dumps({"f": lambda: 0})
    with pytest.raises(t7.T7Error, match="no Torch7 tensor type"):
        t7.dumps(np.zeros(2, dtype=np.complex128))


def test_cache_from_results_feeds_the_bot_cache_protocol(tmp_path):
    from bot7_b200 import bots
    r = np.random.default_rng(1)
    res = {"best": {"x": r.random((1, 2)), "y": np.array([[0.1]]), "t": 3}, "x": r.random((5, 2)), "y": r.random(5)}
    p = tmp_path / "demo_bayesopt.t7"
    t7.save(p, res)
    cache = bots.cache_from_results(str(p), candidates=r.random((50, 2)))
    assert cache["observed"].shape == (5, 2) and cache["responses"].shape == (5, 1) and cache["candidates"].shape == (50, 2)
    assert np.array_equal(cache["observed"], res["x"]) and np.array_equal(cache["responses"][:, 0], res["y"])
    with pytest.raises(ValueError, match="no observations"):
        bots.cache_from_results({"best": {}})
    with pytest.raises(ValueError, match="5 observed points but 4 responses"):
        bots.cache_from_results({"x": res["x"], "y": res["y"][:4]})


def test_bot_save_writes_the_reference_file_name_and_layout(tmp_path, monkeypatch):
    # bots/abstract.lua:234-240: torch.save('demo_' .. torch.type(self) .. '.t7', {best=, x=, y=}); no device needed
    from bot7_b200 import bots
    r = np.random.default_rng(2)
    bot = object.__new__(bots.random_search)
    bot.observed, bot.responses = r.random((4, 3)), r.random((4, 1))
    bot.best = {"x": bot.observed[2:3], "y": bot.responses[2:3], "t": 3}
    monkeypatch.chdir(tmp_path)
    path = bot.save()
    assert path == "demo_bot7.bots.random_search.t7" and (tmp_path / path).exists()
    res = t7.load(str(tmp_path / path))
    assert list(res) == ["best", "x", "y"] and list(res["best"]) == ["x", "y", "t"] and res["best"]["t"] == 3.0
    assert np.array_equal(res["x"], bot.observed) and np.array_equal(res["y"], bot.responses)
    cache = bots.cache_from_results(res)
    assert np.array_equal(cache["observed"], bot.observed) and cache["responses"].shape == (4, 1)
