"""Torch7 binary serialization (`torch.save` / `torch.load`, default binary mode) and the import of a serialized
multi-frame PWC model (back2future.lua:97-129, README.md:49-71).

The pretrained checkpoints (`RoamingImages_H.t7`, ...) are `torch.save`d `nn.gModule`s (train.lua:183), possibly
wrapped in `nn.DataParallelTable` (back2future.lua:113-116).  The file format is Torch7's (torch/File.lua,
`writeObject` / `readObject`), restated from its published definition:

    object   := int32 type, payload
    type 0 nil | 1 number: float64 | 2 string: int32 n, n bytes | 5 boolean: int32
    type 3 table:  int32 index; first occurrence: int32 n, n x (key object, value object)
    type 4 torch:  int32 index; first occurrence: string version ("V 1"), string class name, class payload
                   torch.*Tensor : int32 ndim, ndim x int64 size, ndim x int64 stride, int64 storageOffset (1-based),
                                   storage object
                   torch.*Storage: int64 n, n raw elements
                   any other class (nn.*, nngraph.*, cudnn.*): one table object holding its fields
    type 6 / 7 / 8 function: int32 index; first occurrence: int32 n, n bytes of bytecode, upvalue object

Everything is little-endian; an `index` seen before refers to the object read earlier (this is how the siamese
clones' shared weights and nngraph's node / data tables keep their identity).

None of the checkpoints is available offline: the reader is tested against files written by `save` below, i.e.
against this module's own understanding of the format and of nngraph's object layout -- stated as such in DESIGN.md.
"""
from __future__ import annotations

import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_STORAGE_DTYPES = {
    "torch.FloatStorage": np.float32, "torch.CudaStorage": np.float32, "torch.DoubleStorage": np.float64,
    "torch.CudaDoubleStorage": np.float64, "torch.LongStorage": np.int64, "torch.CudaLongStorage": np.int64,
    "torch.IntStorage": np.int32, "torch.ShortStorage": np.int16, "torch.ByteStorage": np.uint8,
    "torch.CudaByteStorage": np.uint8, "torch.CharStorage": np.int8, "torch.HalfStorage": np.float16,
}
_TENSOR_STORAGE = {k.replace("Storage", "Tensor"): k for k in _STORAGE_DTYPES}


class TorchObject:
    """A deserialized instance of a Torch7 class without a custom reader: `typename` + its field table."""

    def __init__(self, typename, fields=None):
        self.typename = typename
        self.fields = fields if fields is not None else {}

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return "<%s>" % self.typename


class LuaFunction:
    def __init__(self, code, upvalues):
        self.code, self.upvalues = code, upvalues


def lua_list(tbl):
    """The array part {t[1], t[2], ...} of a deserialized Lua table (keys arrive as floats)."""
    out, i = [], 1
    while i in tbl:
        out.append(tbl[i])
        i += 1
    return out


# -------------------------------------------------------------------------------------------------------------
# reader
# -------------------------------------------------------------------------------------------------------------

class _Reader:
    def __init__(self, data):
        self.d, self.p, self.seen = data, 0, {}

    def _take(self, n):
        if self.p + n > len(self.d):
            raise ValueError("t7: truncated file (wanted %d bytes at offset %d of %d)" % (n, self.p, len(self.d)))
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def i32(self):
        return struct.unpack("<i", self._take(4))[0]

    def i64(self):
        return struct.unpack("<q", self._take(8))[0]

    def string(self):
        n = self.i32()
        if n < 0:
            raise ValueError("t7: negative string length at offset %d" % self.p)
        return self._take(n).decode("latin1")

    def obj(self):
        t = self.i32()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = struct.unpack("<d", self._take(8))[0]
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.i32() == 1
        if t == TYPE_TABLE:
            idx = self.i32()
            if idx in self.seen:
                return self.seen[idx]
            tbl = {}
            self.seen[idx] = tbl
            n = self.i32()
            for _ in range(n):
                k = self.obj()
                v = self.obj()
                if isinstance(k, (dict, list, np.ndarray)):
                    k = ("#id", id(k))             # nngraph's mapindex is also keyed by tables
                tbl[k] = v
            return tbl
        if t == TYPE_TORCH:
            idx = self.i32()
            if idx in self.seen:
                return self.seen[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version
            if cls in _TENSOR_STORAGE:
                nd = self.i32()
                size = [self.i64() for _ in range(nd)]
                stride = [self.i64() for _ in range(nd)]
                off = self.i64() - 1
                holder = [None]
                self.seen[idx] = holder
                st = self.obj()
                if st is None or nd == 0:
                    arr = np.zeros(size if nd else (0,), _STORAGE_DTYPES[_TENSOR_STORAGE[cls]])
                else:
                    arr = np.lib.stride_tricks.as_strided(st[off:], shape=size,
                                                          strides=[s * st.itemsize for s in stride])
                self.seen[idx] = arr
                return arr
            if cls in _STORAGE_DTYPES:
                n = self.i64()
                dt = np.dtype(_STORAGE_DTYPES[cls])
                arr = np.frombuffer(self._take(n * dt.itemsize), dt).copy()
                self.seen[idx] = arr
                return arr
            o = TorchObject(cls)
            self.seen[idx] = o
            f = self.obj()
            o.fields = f if isinstance(f, dict) else {"__value": f}
            return o
        if t in (TYPE_FUNCTION, TYPE_RECUR_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION):
            idx = self.i32()
            if idx in self.seen:
                return self.seen[idx]
            fn = LuaFunction(None, None)
            self.seen[idx] = fn
            fn.code = self._take(self.i32())
            fn.upvalues = self.obj()
            return fn
        raise ValueError("t7: unknown type tag %d at offset %d" % (t, self.p - 4))


def load(path):
    """`torch.load(path)`: numbers -> int / float, strings, booleans, tables -> dict, tensors -> numpy arrays (views
    of their storages, strides honoured), other torch classes -> TorchObject."""
    with open(path, "rb") as f:
        return _Reader(f.read()).obj()


# -------------------------------------------------------------------------------------------------------------
# writer (enough of `torch.save` to round-trip models of this family; used by export_model and the tests)
# -------------------------------------------------------------------------------------------------------------

class _Writer:
    def __init__(self):
        self.out, self.index, self.next = [], {}, 1
        self.keep = []

    def i32(self, v):
        self.out.append(struct.pack("<i", v))

    def i64(self, v):
        self.out.append(struct.pack("<q", v))

    def string(self, s):
        b = s.encode("latin1")
        self.i32(len(b))
        self.out.append(b)

    def _ref(self, o):
        """(index, first occurrence?)"""
        k = id(o)
        if k in self.index:
            return self.index[k], False
        self.index[k] = self.next
        self.keep.append(o)
        self.next += 1
        return self.index[k], True

    def obj(self, o):
        if o is None:
            self.i32(TYPE_NIL)
        elif isinstance(o, bool):
            self.i32(TYPE_BOOLEAN)
            self.i32(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.i32(TYPE_NUMBER)
            self.out.append(struct.pack("<d", float(o)))
        elif isinstance(o, str):
            self.i32(TYPE_STRING)
            self.string(o)
        elif isinstance(o, (dict, list, tuple)):
            self.i32(TYPE_TABLE)
            idx, new = self._ref(o)
            self.i32(idx)
            if new:
                items = list(o.items()) if isinstance(o, dict) else [(i + 1, v) for i, v in enumerate(o)]
                self.i32(len(items))
                for k, v in items:
                    self.obj(k)
                    self.obj(v)
        elif isinstance(o, np.ndarray):
            names = {np.dtype(np.float32): "torch.FloatTensor", np.dtype(np.float64): "torch.DoubleTensor",
                     np.dtype(np.int64): "torch.LongTensor", np.dtype(np.uint8): "torch.ByteTensor",
                     np.dtype(np.int32): "torch.IntTensor"}
            cls = names[o.dtype]
            self.i32(TYPE_TORCH)
            idx, new = self._ref(o)
            self.i32(idx)
            if new:
                self.string("V 1")
                self.string(cls)
                a = np.ascontiguousarray(o)
                self.i32(a.ndim)
                for s in a.shape:
                    self.i64(s)
                for s in a.strides:
                    self.i64(s // a.itemsize)
                self.i64(1)
                self.i32(TYPE_TORCH)
                self.i32(self.next)
                self.next += 1
                self.string("V 1")
                self.string(_TENSOR_STORAGE[cls])
                self.i64(a.size)
                self.out.append(a.tobytes())
        elif isinstance(o, TorchObject):
            self.i32(TYPE_TORCH)
            idx, new = self._ref(o)
            self.i32(idx)
            if new:
                self.string("V 1")
                self.string(o.typename)
                self.obj(o.fields)
        else:
            raise TypeError("t7.save: cannot serialize %r" % type(o))


def save(path, obj):
    w = _Writer()
    w.obj(obj)
    with open(path, "wb") as f:
        f.write(b"".join(w.out))


# -------------------------------------------------------------------------------------------------------------
# model import / export
# -------------------------------------------------------------------------------------------------------------

_CONV = ("nn.SpatialConvolution", "cudnn.SpatialConvolution", "nn.SpatialConvolutionMM")
_LEVEL_OF_WIDTH = {16: 2, 32: 3, 64: 4, 96: 5, 128: 6, 192: 7}


def _unwrap(model):
    """back2future.lua:113-116: a DataParallelTable holds the replica in modules[1]."""
    if isinstance(model, TorchObject) and model.typename == "nn.DataParallelTable":
        model = lua_list(model["modules"])[0]
    if not isinstance(model, TorchObject) or model.typename != "nn.gModule":
        raise ValueError("t7: expected an nn.gModule (or a DataParallelTable of one), got %r" % (model,))
    return model


def import_model(model, win=9):
    """Serialized `createModelMulti` graph -> (params dict with this repo's names, past_flow).

    The graph is identified structurally, not by node order (nngraph's topological order is not the creation order):
      * an nn.Sequential with two convolutions is a convUnit; its level follows from nOutputPlane (16 -> l2 ... 192 -> l7);
      * an nn.Sequential with six convolutions is a decoder; its level follows from the first nInputPlane
        (2 win^2 [+ C_l + 2]); it is the occlusion decoder if an nn.SpatialSoftMax consumes it, otherwise a flow
        decoder -- the FUTURE one if a positive nn.MulConstant (pwc.lua:404, 443) is reachable through
        SpatialUpSamplingBilinear nodes only, else the PAST one."""
    g = _unwrap(model)
    nodes = lua_list(g["forwardnodes"])
    nd = 2 * win * win

    def module_of(node):
        data = node["data"] if isinstance(node, TorchObject) else node.get("data")
        return data.get("module") if isinstance(data, dict) else None

    def children(node):
        ch = node["children"] if isinstance(node, TorchObject) else node.get("children")
        return lua_list(ch) if isinstance(ch, dict) else list(ch or [])

    def convs_of(seq):
        return [m for m in lua_list(seq["modules"]) if isinstance(m, TorchObject) and m.typename in _CONV]

    def reaches_positive_mul(node, depth=0):
        for c in children(node):
            m = module_of(c)
            if m is None:
                continue
            if m.typename == "nn.MulConstant" and float(m["constant_scalar"]) > 0:
                return True
            if m.typename == "nn.SpatialUpSamplingBilinear" and depth < 8 and reaches_positive_mul(c, depth + 1):
                return True
        return False

    params, past_flow = {}, False

    def put(prefix, convs):
        for i, c in enumerate(convs):
            w = np.asarray(c["weight"], np.float32)
            n_out, n_in = int(c["nOutputPlane"]), int(c["nInputPlane"])
            params["%s.%d.weight" % (prefix, i)] = np.ascontiguousarray(w.reshape(n_out, n_in, 3, 3))
            params["%s.%d.bias" % (prefix, i)] = np.ascontiguousarray(np.asarray(c["bias"], np.float32).reshape(n_out))

    for node in nodes:
        m = module_of(node)
        if not isinstance(m, TorchObject) or m.typename != "nn.Sequential":
            continue
        convs = convs_of(m)
        if len(convs) == 2:
            l = _LEVEL_OF_WIDTH.get(int(convs[0]["nOutputPlane"]))
            if l is None:
                raise ValueError("t7: convUnit with %d output planes" % int(convs[0]["nOutputPlane"]))
            put("feat.l%d" % l, convs)
        elif len(convs) == 6:
            n_in = int(convs[0]["nInputPlane"])
            level = None
            for width, l in _LEVEL_OF_WIDTH.items():
                if n_in in (nd + width + 2, nd + width) and l >= 3:
                    level = l
            if n_in == nd:
                level = 7
            if level is None:
                raise ValueError("t7: decoder with %d input planes does not belong to this model family" % n_in)
            is_occ = any(getattr(module_of(c), "typename", "") == "nn.SpatialSoftMax" for c in children(node))
            if is_occ:
                kind = "occ"
            elif reaches_positive_mul(node):
                kind = "flow"
            else:
                kind, past_flow = "bflow", True
            put("%s.l%d" % (kind, level), convs)
    return params, past_flow


def export_model(params, past_flow=False, win=9, levels=7, l_st=3, flownet_factor=20):
    """The object tree `torch.save` would write for createModelMulti's gModule, reduced to what `import_model` and
    back2future.lua:init read: forwardnodes with data.module / children, Sequential containers, convolution fields,
    the MulConstant / up-sampling / SoftMax consumers of the decoders.  For tests and for handing weights back."""
    feat = (3, 16, 32, 64, 96, 128, 192)
    nd = 2 * win * win

    def conv(name, stride):
        w = np.asarray(params[name + ".weight"], np.float32)
        return TorchObject("nn.SpatialConvolution", {
            "weight": w, "bias": np.asarray(params[name + ".bias"], np.float32), "nOutputPlane": w.shape[0],
            "nInputPlane": w.shape[1], "kW": 3, "kH": 3, "dW": stride, "dH": stride, "padW": 1, "padH": 1})

    def node(module, kids=()):
        return TorchObject("nngraph.Node", {"data": {"module": module}, "children": list(kids)})

    nodes = []
    units = {}
    for l in range(2, levels + 1):
        units[l] = TorchObject("nn.Sequential", {"modules": [
            conv("feat.l%d.0" % l, 2), TorchObject("nn.LeakyReLU", {"negval": 0.2}),
            conv("feat.l%d.1" % l, 1), TorchObject("nn.LeakyReLU", {"negval": 0.2})]})
    for _f in range(3):                    # three siamese towers sharing the module objects' tensors
        for l in range(2, levels + 1):
            nodes.append(node(units[l]))

    def decoder(kind, l):
        mods = []
        for i in range(6):
            mods.append(conv("%s.l%d.%d" % (kind, l, i), 1))
            if i < 5:
                mods.append(TorchObject("nn.LeakyReLU", {"negval": 0.2}))
        return TorchObject("nn.Sequential", {"modules": mods})

    for l in range(levels, l_st - 1, -1):
        sm = node(TorchObject("nn.SpatialSoftMax", {}))
        nodes += [node(decoder("occ", l), [sm]), sm]
        for kind, sgns in (("flow", (1.0,) if past_flow else (-1.0, 1.0)),) + ((("bflow", (-1.0,)),) if past_flow else ()):
            muls = [node(TorchObject("nn.MulConstant", {"constant_scalar": flownet_factor * s / 2.0 ** (l - l_st)}))
                    for s in sgns]
            up2 = node(TorchObject("nn.SpatialUpSamplingBilinear", {"scale_factor": 2}), muls)
            up1 = node(TorchObject("nn.SpatialUpSamplingBilinear", {"scale_factor": 2}), [up2])
            nodes += [node(decoder(kind, l), [up1]), up1, up2] + muls
    return TorchObject("nn.gModule", {"forwardnodes": nodes, "flow_scale": [flownet_factor / 2.0 ** (l - l_st)
                                                                            for l in range(levels, l_st - 1, -1)],
                                      "past_flow": bool(past_flow)})
