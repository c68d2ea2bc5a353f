"""Host side of computeFlow / init: image.scale restatement, Torch7 serialization, model import."""
import numpy as np
import pytest

from back2future_b200 import imageio as io, t7
from oracle import pwc_oracle as po


def test_scale_identity_and_shapes():
    rng = np.random.default_rng(0)
    a = rng.random((9, 37, 53)).astype(np.float32)
    assert np.array_equal(io.scale(a, 53, 37), a)
    assert io.scale(a, 64, 32).shape == (9, 32, 64)
    assert io.scale(a[0], 20, 10).shape == (10, 20)
    with pytest.raises(ValueError):
        io.scale(a, 10, 10, "bicubic")


def test_scale_bilinear_enlarging_is_align_corners():
    a = np.arange(12, dtype=np.float32).reshape(1, 3, 4)
    up = io.scale(a, 7, 5)
    np.testing.assert_allclose(up[0, 0], np.arange(7) * 0.5, rtol=1e-6)
    np.testing.assert_allclose(up[0, :, 0], np.arange(5) * 2.0, rtol=1e-6)
    assert up[0, -1, -1] == 11.0


def test_scale_bilinear_shrinking_is_a_box_average():
    """image.c scaleLinear_rowcol, dst < src: fractional-weight average of the covered source samples; a constant
    image stays constant, an integer factor is the plain block mean, and the mean is preserved to fp32 accuracy."""
    rng = np.random.default_rng(1)
    c = np.full((1, 10, 30), 3.25, np.float32)
    np.testing.assert_allclose(io.scale(c, 7, 3), 3.25, rtol=1e-6)
    a = rng.random((2, 8, 12)).astype(np.float32)
    blk = a.reshape(2, 4, 2, 4, 3).mean(axis=(2, 4))
    np.testing.assert_allclose(io.scale(a, 4, 4), blk, rtol=1e-5)
    big = rng.random((3, 375, 1242)).astype(np.float32)
    sm = io.scale(big, 1216, 320)
    assert abs(sm.mean() - big.mean()) < 1e-3
    # one row by hand: src 5 -> dst 2, scale 2.5: [x0 + x1 + .5 x2] / 2.5, [.5 x2 + x3 + x4] / 2.5
    r = np.array([[[1, 2, 3, 4, 5]]], np.float32)
    np.testing.assert_allclose(io.scale(r, 2, 1)[0, 0], [(1 + 2 + 1.5) / 2.5, (1.5 + 4 + 5) / 2.5], rtol=1e-6)


def test_scale_simple_is_nearest_floor():
    a = np.arange(20, dtype=np.float64).reshape(1, 4, 5)
    s = io.scale(a, 10, 8, "simple")
    assert s.dtype == np.float64 and s.shape == (1, 8, 10)
    assert np.array_equal(s[0, :, 0], [0, 0, 5, 5, 10, 10, 15, 15])
    assert np.array_equal(s[0, 0], [0, 0, 1, 1, 2, 2, 3, 3, 4, 4])
    m = (a > 7).astype(np.uint8)
    assert io.scale(m, 3, 2, "simple").dtype == np.uint8
    assert np.array_equal(io.scale(a, 5, 4, "simple"), a)


def test_t7_round_trip(tmp_path):
    shared = np.arange(6, dtype=np.float32).reshape(2, 3)
    obj = {"a": 1.5, "b": "text", "c": True, "list": [1, 2, 3], "t": shared, "t2": shared,
           "mod": t7.TorchObject("nn.Linear", {"weight": np.ones((2, 2), np.float64), "n": 7}),
           "long": np.array([3, 4], np.int64), 5: None}
    p = str(tmp_path / "x.t7")
    t7.save(p, obj)
    back = t7.load(p)
    assert back["a"] == 1.5 and back["b"] == "text" and back["c"] is True
    assert t7.lua_list(back["list"]) == [1, 2, 3]
    assert np.array_equal(back["t"], shared) and back["t"] is back["t2"]        # identity of shared objects survives
    assert back["mod"].typename == "nn.Linear" and back["mod"]["n"] == 7
    assert back["mod"]["weight"].dtype == np.float64
    assert np.array_equal(back["long"], [3, 4])
    with open(p, "rb") as f:
        raw = f.read()
    with pytest.raises(ValueError):
        t7._Reader(raw[:-5]).obj()


def test_t7_known_bytes():
    """A hand-assembled file: the table {1 = 2.5, x = "hi"} as torch.save writes it."""
    import struct
    b = struct.pack("<i", 3) + struct.pack("<i", 1) + struct.pack("<i", 2)
    b += struct.pack("<i", 1) + struct.pack("<d", 1.0) + struct.pack("<i", 1) + struct.pack("<d", 2.5)
    b += struct.pack("<i", 2) + struct.pack("<i", 1) + b"x" + struct.pack("<i", 2) + struct.pack("<i", 2) + b"hi"
    assert t7._Reader(b).obj() == {1: 2.5, "x": "hi"}
    # a FloatTensor of size 2x2 viewing a 6-element storage at offset 2 with strides (2, 1)
    s = lambda x: struct.pack("<i", len(x)) + x
    tb = struct.pack("<i", 4) + struct.pack("<i", 1) + s(b"V 1") + s(b"torch.FloatTensor") + struct.pack("<i", 2)
    tb += struct.pack("<qq", 2, 2) + struct.pack("<qq", 2, 1) + struct.pack("<q", 2)
    tb += struct.pack("<i", 4) + struct.pack("<i", 2) + s(b"V 1") + s(b"torch.FloatStorage") + struct.pack("<q", 6)
    tb += np.arange(6, dtype=np.float32).tobytes()
    assert np.array_equal(t7._Reader(tb).obj(), [[1, 2], [3, 4]])


@pytest.mark.parametrize("past_flow", [False, True])
def test_model_export_import(tmp_path, past_flow):
    opt = po.Opt(past_flow=past_flow)
    params = po.init_params(opt, seed=9)
    p = str(tmp_path / "m.t7")
    model = t7.export_model(params, past_flow)
    wrapped = t7.TorchObject("nn.DataParallelTable", {"modules": [model]})
    t7.save(p, wrapped)
    got, pf = t7.import_model(t7.load(p))
    assert pf == past_flow
    assert set(got) == set(params)
    for k in params:
        assert np.array_equal(got[k], params[k]), k
    with pytest.raises(ValueError):
        t7.import_model(t7.TorchObject("nn.Sequential", {}))


def test_backward_plan_lanes_order_their_dependencies(monkeypatch):
    """pwc._Plan.launch_backward without a GPU: fake streams / events record who waits for what.  A weight gradient
    (lane 1) waits for the lane that produced its operands at that point of the plan; a flow decoder's chain (lane 2)
    starts behind the main stream and its accumulating last call behind the occlusion decoder's write; the first call
    that reads the joined gradient waits for the chains; everything is joined at the end of a slice unless join=False,
    in which case `backward_streams()` names the streams a collective has to wait for."""
    import ctypes as C
    from back2future_b200 import pwc

    log = []

    class FakeStream:
        n = 0

        def __init__(self):
            FakeStream.n += 1
            self.cuda_stream = 1000 + FakeStream.n
            self.issued = 0                       # work items issued so far (what an event recorded now covers)

        def wait_event(self, ev):
            log.append(("wait", self.cuda_stream, ev.stream.cuda_stream, ev.mark))

        def wait_stream(self, other):
            log.append(("join", self.cuda_stream, other.cuda_stream, other.issued))

    class FakeEvent:
        def record(self, stream):
            self.stream, self.mark = stream, stream.issued

    main = FakeStream()
    monkeypatch.setattr(pwc.torch.cuda, "current_stream", lambda *a, **k: main)
    monkeypatch.setattr(pwc.torch.cuda, "Stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(pwc.torch.cuda, "Event", lambda *a, **k: FakeEvent())
    by_handle = {}

    def op(name):
        def fn(stream_handle):
            s = by_handle[stream_handle.value]
            s.issued += 1
            log.append(("run", name, s.cuda_stream))
            return 0
        return fn

    p = pwc._Plan(1, 64, 64)
    p.bops = [
        (op("g_fs"), ()),                              # main: produces the flow decoder's output gradient
        (op("flow.split"), (), 2, (0,)),               # flow chain starts behind it
        (op("flow.wgrad5"), (), 1, (2,)),              # its weight gradient: behind the chain, on lane 1
        (op("flow.dgrad5"), (), 2),
        (op("occ.dgrad0 (write)"), ()),                # occlusion decoder on the main stream
        (op("occ.wgrad0"), (), 1),                     # default dependency: the main stream
        (op("flow.dgrad0 (add)"), (), 2, (0,)),        # behind the occlusion decoder's write
        (op("costvol_bwd"), (), 0, (2, 3)),            # reads the joined gradient: behind the chains (lane 3 unused)
    ]
    # the executor creates its side streams on first use; register them for the fake launch functions
    p.launch_backward(0, 0)
    by_handle[main.cuda_stream] = main
    for s in p._blanes.values():
        by_handle[s.cuda_stream] = s
    lane = {k: s.cuda_stream for k, s in p._blanes.items()}
    lane[0] = main.cuda_stream
    log.clear()
    p.launch_backward(join=False)
    runs = [e for e in log if e[0] == "run"]
    assert [(n, s) for _r, n, s in runs] == [
        ("g_fs", lane[0]), ("flow.split", lane[2]), ("flow.wgrad5", lane[1]), ("flow.dgrad5", lane[2]),
        ("occ.dgrad0 (write)", lane[0]), ("occ.wgrad0", lane[1]), ("flow.dgrad0 (add)", lane[2]), ("costvol_bwd", lane[0])]

    def waits_before(name):
        i = log.index(next(e for e in log if e[0] == "run" and e[1] == name))
        out = []
        while i > 0 and log[i - 1][0] == "wait":
            i -= 1
            out.append(log[i][1:])
        return out

    assert waits_before("flow.split") == [(lane[2], lane[0], 1)]              # behind g_fs
    assert waits_before("flow.wgrad5") == [(lane[1], lane[2], 1)]             # behind flow.split
    assert waits_before("flow.dgrad5") == []                                  # a lane is ordered in itself
    assert waits_before("occ.wgrad0") == [(lane[1], lane[0], 2)]              # behind occ.dgrad0
    assert waits_before("flow.dgrad0 (add)") == [(lane[2], lane[0], 2)]       # behind the occlusion decoder's write
    assert waits_before("costvol_bwd") == [(lane[0], lane[2], 3)]             # lane 3 carries nothing: no wait on it
    assert not [e for e in log if e[0] == "join"]
    assert [s.cuda_stream for s in p.backward_streams()] == [lane[1], lane[2]]
    # the next slice joins what is still in flight
    p.bops = [(op("adam"), ())]
    log.clear()
    p.launch_backward(lo=1, hi=1, join=True)          # lo > 0: a later slice of the same plan run
    assert sorted(e[2] for e in log if e[0] == "join") == sorted([lane[1], lane[2]])
    assert p.backward_streams() == []
    # one stream only: same order, no waits
    monkeypatch.setattr(pwc, "SIDE_LANES", False)
    p.bops = [(op("a"), (), 1, (2,)), (op("b"), (), 2, (0,))]
    log.clear()
    p.launch_backward()
    assert log == [("run", "a", lane[0]), ("run", "b", lane[0])]
