"""`.flo` wire format (flowExtensions.lua:254-287).  Host-side: runs without a GPU.
The expected bytes are built independently with struct/numpy from the published Middlebury layout the reference
follows: float32 tag 202021.25, int32 width, int32 height, interleaved (u, v) float32 rows, little-endian."""
import struct

import numpy as np
import pytest

from back2future_b200 import _lib, flowio


def expected_bytes(F):
    _, h, w = F.shape
    return struct.pack("<fii", 202021.25, w, h) + np.ascontiguousarray(F.transpose(1, 2, 0), "<f4").tobytes()


@pytest.mark.parametrize("h,w", [(1, 1), (3, 5), (37, 64), (33, 1025)])
def test_write_matches_the_reference_layout_and_round_trips(tmp_path, h, w):
    F = np.random.default_rng(h * 1000 + w).standard_normal((2, h, w)).astype(np.float32)
    path = tmp_path / "a.flo"
    flowio.writeFLO(path, F)
    raw = path.read_bytes()
    assert raw == expected_bytes(F)
    assert raw[:4] == b"PIEH"
    back = flowio.loadFLO(path)
    assert back.dtype == np.float32 and back.shape == (2, h, w) and np.array_equal(back, F)


def test_read_a_file_written_by_someone_else(tmp_path):
    F = np.arange(2 * 4 * 6, dtype=np.float32).reshape(2, 4, 6)
    path = tmp_path / "b.flo"
    path.write_bytes(expected_bytes(F))
    assert np.array_equal(flowio.loadFLO(path), F)


def test_errors_like_the_reference(tmp_path):
    bad = tmp_path / "bad.flo"
    bad.write_bytes(struct.pack(">fii", 202021.25, 2, 2) + b"\0" * 32)      # big-endian tag
    with pytest.raises(_lib.B2FError, match="bigendian"):
        flowio.loadFLO(bad)
    short = tmp_path / "short.flo"
    short.write_bytes(expected_bytes(np.zeros((2, 3, 3), np.float32))[:-4])
    with pytest.raises(_lib.B2FError, match="truncated"):
        flowio.loadFLO(short)
    with pytest.raises(_lib.B2FError, match="cannot open"):
        flowio.loadFLO(tmp_path / "missing.flo")
    with pytest.raises(ValueError):
        flowio.writeFLO(tmp_path / "c.flo", np.zeros((3, 4, 4), np.float32))
    lib = _lib.load()
    buf = np.zeros((2, 2, 2), np.float32)
    ok = tmp_path / "ok.flo"
    flowio.writeFLO(ok, np.zeros((2, 3, 3), np.float32))
    assert lib.b2f_flo_read(str(ok).encode(), buf.ctypes.data, 2, 2) != 0      # size mismatch is refused
