"""Host-side pre/post-processing around `computeFlow` (back2future.lua:33-95).

These are the pieces of the reference's inference glue that are plain arithmetic on host tensors -- they run on
the CPU in the reference too (back2future.lua builds the 9-channel input on the host and only then calls
`:cuda()`), so they are host code here as well, not a CPU fallback of a device kernel:

* `load_png`        -- what `image.load(path, 3, 'float')` hands to `computeFlow` (README.md:49-60): 8-bit PNG ->
                       float32 (3, H, W) in [0, 1].  Own decoder (zlib from the standard library + numpy).
* `color_normalize` -- `TF.ColorNormalize(meanstd)` (transforms.lua:33-45) with back2future.lua:33-36's constants,
                       applied to every RGB triple of the (3k, H, W) stack.
* `fine_size`       -- back2future.lua:55-67: width / height rounded DOWN to a multiple of 64 (7 pyramid levels).
* `occlusion_masks` -- back2future.lua:87-92: threshold 0.6666 evaluated in DOUBLE (quirk Q13), future occlusions
                       from channel 2, past from channel 1 of the occlusion map.
* `rescale_flow`    -- back2future.lua:80-84: u, v multiplied by width / height ratios after the resize.

* `scale`           -- `image.scale(src, width, height [, mode])` (back2future.lua:71, 82, 89-91).  The Torch7 `image`
                       package is not part of the reference tree; its two modes are restated from the package's C
                       source as published (generic/image.c: `scaleBilinear` = separable `scaleLinear_rowcol`, rows
                       then columns through a temporary, align-corners interpolation when enlarging and a box average
                       with fractional end weights when shrinking, float arithmetic; `scaleSimple` = nearest sample at
                       floor(i * src / dst)).  Unpinned: no Torch7 here to run it against.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

MEAN = (0.485, 0.456, 0.406)   # back2future.lua:33-36
STD = (0.229, 0.224, 0.225)
OCC_THRESHOLD = 0.6666         # back2future.lua:40

_PNG_MAGIC = b"\x89PNG\r\n\x1a\n"


def _unfilter(raw, h, stride, bpp):
    """PNG scanline filters (0 None, 1 Sub, 2 Up, 3 Average, 4 Paeth), byte-wise mod 256."""
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    pos = 0
    for y in range(h):
        ft = raw[pos]
        line = np.frombuffer(raw, np.uint8, stride, pos + 1).astype(np.int32)
        pos += 1 + stride
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:
            # out[i] = line[i] + out[i - bpp]: a running sum per byte lane
            cur = line.copy()
            lanes = cur[: (stride // bpp) * bpp].reshape(-1, bpp)
            lanes[:] = np.cumsum(lanes, axis=0) & 255
            if stride % bpp:
                raise ValueError("load_png: scanline length is not a multiple of the pixel size")
        elif ft in (3, 4):
            cur = np.zeros(stride, np.int32)
            ln, pv = line.tolist(), prev.tolist()
            cl = [0] * stride
            for i in range(stride):
                a = cl[i - bpp] if i >= bpp else 0
                b = pv[i]
                if ft == 3:
                    pred = (a + b) >> 1
                else:
                    c = pv[i - bpp] if i >= bpp else 0
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cl[i] = (ln[i] + pred) & 255
            cur = np.asarray(cl, np.int32)
        else:
            raise ValueError("load_png: unknown filter type %d" % ft)
        out[y] = cur.astype(np.uint8)
        prev = cur
    return out


def load_png(path):
    """8-bit grey / RGB / RGBA / grey+alpha, non-interlaced PNG -> float32 (3, H, W) in [0, 1] (alpha dropped, grey
    replicated), i.e. `image.load(path, 3, 'float')`.  Anything else raises ValueError."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != _PNG_MAGIC:
        raise ValueError("load_png: %s is not a PNG file" % path)
    pos, idat, hdr = 8, [], None
    while pos + 8 <= len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if len(body) != n:
            raise ValueError("load_png: truncated chunk")
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        if zlib.crc32(typ + body) & 0xFFFFFFFF != crc:
            raise ValueError("load_png: bad CRC in %s chunk" % typ.decode("latin1"))
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        pos += 12 + n
    if hdr is None or not idat:
        raise ValueError("load_png: missing IHDR / IDAT")
    w, h, depth, ctype, comp, flt, interlace = hdr
    nch = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if depth != 8 or nch is None or comp != 0 or flt != 0 or interlace != 0:
        raise ValueError("load_png: only 8-bit non-interlaced grey / RGB (with or without alpha) is supported "
                         "(depth %d, colour type %d, interlace %d)" % (depth, ctype, interlace))
    raw = zlib.decompress(b"".join(idat))
    stride = w * nch
    if len(raw) != h * (stride + 1):
        raise ValueError("load_png: image data has %d bytes, expected %d" % (len(raw), h * (stride + 1)))
    px = _unfilter(raw, h, stride, nch).reshape(h, w, nch)
    rgb = px[:, :, :3] if nch >= 3 else np.repeat(px[:, :, :1], 3, axis=2)
    return np.ascontiguousarray(rgb.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)


def color_normalize(imgs, mean=MEAN, std=STD):
    """transforms.lua:33-45: for every RGB triple c of a (3k, H, W) stack, channel i: (x - mean[i]) / std[i], as two
    separately rounded fp32 operations (`add(-mean)` then `div(std)`); returns a copy."""
    imgs = np.array(imgs, dtype=np.float32, copy=True)
    if imgs.ndim != 3 or imgs.shape[0] % 3:
        raise ValueError("color_normalize: expected a (3k, H, W) array, got %r" % (imgs.shape,))
    for c in range(imgs.shape[0] // 3):
        for i in range(3):
            ch = imgs[3 * c + i]
            ch += np.float32(-mean[i])
            ch /= np.float32(std[i])
    return imgs


def fine_size(width, height):
    """back2future.lua:55-67: the network's input size -- both edges rounded down to a multiple of 64."""
    return width - width % 64, height - height % 64


def _scale_linear_axis(src, dst_len):
    """image.c `scaleLinear_rowcol` along the LAST axis of a float32 array, all other axes in parallel."""
    src_len = src.shape[-1]
    f32 = np.float32
    if dst_len == src_len:
        return src.copy()
    dst = np.empty(src.shape[:-1] + (dst_len,), np.float32)
    if dst_len > src_len:
        if src_len == 1:
            dst[...] = src[..., :1]
            return dst
        sc = f32(src_len - 1) / f32(dst_len - 1)
        di = np.arange(dst_len - 1)
        si_f = (di.astype(np.float32) * sc).astype(np.float32)
        si_i = si_f.astype(np.int64)
        fr = (si_f - si_i.astype(np.float32)).astype(np.float32)
        dst[..., :-1] = (f32(1) - fr) * src[..., si_i] + fr * src[..., si_i + 1]
        dst[..., -1] = src[..., -1]
        return dst
    sc = f32(src_len) / f32(dst_len)
    si0_i, si0_f = 0, f32(0)
    for di in range(dst_len):
        si1_f = f32(di + 1) * sc
        si1_i = int(si1_f)
        si1_f = f32(si1_f - f32(si1_i))
        acc = (f32(1) - si0_f) * src[..., si0_i]
        n = f32(1) - si0_f
        for si in range(si0_i + 1, si1_i):
            acc = acc + src[..., si]
            n = f32(n + f32(1))
        if si1_i < src_len:
            acc = acc + si1_f * src[..., si1_i]
            n = f32(n + si1_f)
        dst[..., di] = acc / n
        si0_i, si0_f = si1_i, si1_f
    return dst


def scale(src, width, height, mode="bilinear"):
    """`image.scale(src, width, height, mode)` for a (k, H, W) or (H, W) host array.

    'bilinear' (the default, back2future.lua:71): rows first into a (H, width) temporary, then columns -- float32
    throughout, like the package's FloatTensor path.  'simple' (back2future.lua:82, 89-91): dst[j, i] =
    src[min(floor(j * H / height), H - 1), min(floor(i * W / width), W - 1)] with the ratios in float32; the dtype
    is kept (the reference calls it on DoubleTensors and on the ByteTensor masks)."""
    a = np.asarray(src)
    if a.ndim not in (2, 3):
        raise ValueError("scale: expected a (k, H, W) or (H, W) array, got %r" % (a.shape,))
    if width <= 0 or height <= 0:
        raise ValueError("scale: bad target size %dx%d" % (width, height))
    H, W = a.shape[-2], a.shape[-1]
    if mode == "simple":
        scx = np.float32(W) / np.float32(width)
        scy = np.float32(H) / np.float32(height)
        ii = np.minimum((np.arange(width, dtype=np.float32) * scx).astype(np.int64), W - 1)
        jj = np.minimum((np.arange(height, dtype=np.float32) * scy).astype(np.int64), H - 1)
        return np.ascontiguousarray(a[..., jj[:, None], ii[None, :]])
    if mode != "bilinear":
        raise ValueError("scale: mode %r is not built (the reference uses 'bilinear' and 'simple')" % (mode,))
    a = np.ascontiguousarray(a, dtype=np.float32)
    tmp = _scale_linear_axis(a, width)                                         # rows: (.., H, W) -> (.., H, width)
    out = _scale_linear_axis(np.ascontiguousarray(np.swapaxes(tmp, -1, -2)), height)   # columns
    return np.ascontiguousarray(np.swapaxes(out, -1, -2))


def occlusion_masks(occ, threshold=OCC_THRESHOLD):
    """back2future.lua:87-92 on the (2, h, w) occlusion map: converted to double FIRST (Q13), then `>= 0.6666`;
    returns (future, past) = (channel 2, channel 1 of the 1-based Lua indexing) as uint8 (1, h, w) arrays.  Which
    entry of the network's output table is the occlusion map is the caller's business (quirk Q12)."""
    occ = np.asarray(occ)
    if occ.ndim != 3 or occ.shape[0] != 2:
        raise ValueError("occlusion_masks: expected a (2, h, w) array, got %r" % (occ.shape,))
    o = occ.astype(np.float64)
    return (o[1:2] >= threshold).astype(np.uint8), (o[0:1] >= threshold).astype(np.uint8)


def rescale_flow(flow, width, height):
    """back2future.lua:77-84, the arithmetic half: `flow` is the network-size (2, h_net, w_net) flow; u is multiplied
    by width / w_net and v by height / h_net in double (`:double()` at :77).  The reference resizes with the 'simple'
    (nearest) mode first and multiplies afterwards; the two commute."""
    flow = np.asarray(flow, dtype=np.float64).copy()
    if flow.ndim != 3 or flow.shape[0] != 2:
        raise ValueError("rescale_flow: expected a (2, h, w) array, got %r" % (flow.shape,))
    flow[1] *= height / flow.shape[1]
    flow[0] *= width / flow.shape[2]
    return flow
