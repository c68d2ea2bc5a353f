"""Host-side pre/post-processing of computeFlow (back2future_b200/imageio.py)."""
import os

import numpy as np
import pytest

from back2future_b200 import imageio as bio

REF_SAMPLES = "/root/reference/samples"


def _pil():
    return pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("mode", ["RGB", "L", "RGBA", "LA"])
@pytest.mark.parametrize("shape", [(37, 53), (1, 1), (8, 64)])
def test_load_png_against_pil(tmp_path, mode, shape):
    Image = _pil()
    rng = np.random.default_rng(11)
    h, w = shape
    nch = {"RGB": 3, "L": 1, "RGBA": 4, "LA": 2}[mode]
    # smooth + noise, so that the encoder picks different scanline filters
    base = (np.linspace(0, 255, w)[None, :, None] + 40 * rng.standard_normal((h, w, nch))).clip(0, 255).astype(np.uint8)
    arr = base[:, :, 0] if nch == 1 else base
    path = str(tmp_path / "t.png")
    Image.fromarray(arr, mode).save(path)
    got = bio.load_png(path)
    want = np.asarray(Image.open(path).convert("RGB"), np.float32).transpose(2, 0, 1) / np.float32(255)
    if mode in ("RGBA", "LA"):   # PIL's convert drops alpha as well
        want = np.asarray(Image.open(path), np.uint8).reshape(h, w, nch)[:, :, : (3 if mode == "RGBA" else 1)]
        want = (np.repeat(want, 3, axis=2) if mode == "LA" else want).astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    assert got.dtype == np.float32 and got.shape == (3, h, w)
    assert np.array_equal(got, want)


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="reference tree not mounted")
def test_load_png_reference_samples():
    Image = _pil()
    name = "frame_0010.png"
    got = bio.load_png(os.path.join(REF_SAMPLES, name))
    want = np.asarray(Image.open(os.path.join(REF_SAMPLES, name)).convert("RGB"), np.float32).transpose(2, 0, 1) / np.float32(255)
    assert got.shape == want.shape == (3, 375, 1242)      # BASELINE configs[0]: 1242 x 375
    assert np.array_equal(got, want)
    assert bio.fine_size(1242, 375) == (1216, 320)        # the size computeFlow feeds the network (back2future.lua:55-67)


def test_load_png_rejects_what_it_does_not_decode(tmp_path):
    Image = _pil()
    p = str(tmp_path / "p.png")
    Image.fromarray(np.zeros((4, 4), np.uint16)).save(p)      # 16-bit
    with pytest.raises(ValueError):
        bio.load_png(p)
    q = str(tmp_path / "q.png")
    open(q, "wb").write(b"not a png")
    with pytest.raises(ValueError):
        bio.load_png(q)
    good = str(tmp_path / "g.png")
    Image.fromarray(np.zeros((4, 4, 3), np.uint8)).save(good)
    raw = bytearray(open(good, "rb").read())
    raw[40] ^= 0xFF                                          # corrupt a chunk -> CRC mismatch
    open(good, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        bio.load_png(good)


def test_color_normalize_matches_transforms_lua():
    rng = np.random.default_rng(3)
    x = rng.random((9, 5, 7), dtype=np.float32)
    y = bio.color_normalize(x)
    assert y is not x and y.dtype == np.float32
    for c in range(3):
        for i in range(3):
            want = (x[3 * c + i] + np.float32(-bio.MEAN[i])) / np.float32(bio.STD[i])
            assert np.array_equal(y[3 * c + i], want)
    assert abs(float(y.min())) < 2.2 and float(y.max()) < 2.7     # the ColorNormalize range SURVEY 8d quotes
    with pytest.raises(ValueError):
        bio.color_normalize(np.zeros((4, 2, 2), np.float32))


def test_fine_size_and_flow_rescale():
    assert bio.fine_size(1024, 436) == (1024, 384)       # BASELINE configs[4]
    assert bio.fine_size(1024, 448) == (1024, 448)
    f = np.ones((2, 320, 1216), np.float32)
    g = bio.rescale_flow(f, 1242, 375)
    assert g.dtype == np.float64
    assert np.allclose(g[0], 1242 / 1216) and np.allclose(g[1], 375 / 320)


def test_occlusion_threshold_is_evaluated_in_double():
    """Q13: float32(0.6666) is below the double 0.6666, so the reference's `:double()` then `ge` leaves it unmasked."""
    occ = np.zeros((2, 1, 3), np.float32)
    occ[1, 0] = [0.6666, np.nextafter(np.float32(0.6666), np.float32(1)), 0.9]
    occ[0, 0] = [0.1, 0.6667, 0.6666]
    fut, past = bio.occlusion_masks(occ)
    assert float(np.float32(0.6666)) < 0.6666
    assert fut.tolist() == [[[0, 1, 1]]] and past.tolist() == [[[0, 1, 0]]]
    assert fut.dtype == np.uint8 and fut.shape == (1, 1, 3)
