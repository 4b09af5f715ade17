"""ONNX -> VTW1 weight importer, without the `onnx` package.

The reference loads its network from a model file handed to `VitTrack::new(model_path)`
(/root/reference/src/tracker_context.rs:21, path constant /root/reference/src/main.rs:25); the
public form of that network is an ONNX file with inputs `template[1,3,128,128]`,
`search[1,3,256,256]` and three 16x16 maps out.  This module reads such a
file with a small protobuf wire-format decoder, detects the architecture (D, depth, heads,
hidden, head channels) from the graph, and writes the flat VTW1 file that the CPU oracle and
the CUDA path load (weights.py).

What is pinned and what is not: the importer round-trips the stand-in network's own ONNX
export (tools/torch_model.py) bit-exactly back to the VTW1 file it was exported from
(tests/test_onnx_import.py).  The real `object_tracking_vittrack_2023sep.onnx` is not
available offline and its topology is unknown; a graph that does not have the declared
one-stream structure (one 16x16 patch conv, `depth` x [LayerNorm, QKV MatMul, Softmax, proj,
LayerNorm, FC1, GELU (Erf), FC2], final LayerNorm, 3x3 conv + ReLU + 1x1 conv head) is
rejected with a message that names what was found instead of being force-fitted.

Host-side tooling only: nothing here runs on the per-frame path.
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import weights as W


class OnnxImportError(ValueError):
    pass


# ---- protobuf wire format ---------------------------------------------------------------------------------------------------

def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        if pos >= len(buf):
            raise OnnxImportError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise OnnxImportError("varint too long")


def _fields(buf: memoryview):
    """Yields (field_number, wire_type, value) — value is an int (varint / fixed) or a memoryview (length-delimited)."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise OnnxImportError("truncated length-delimited field")
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise OnnxImportError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [v]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _signed(x: int) -> int:
    return x - (1 << 64) if x >= 1 << 63 else x


# ---- the parts of onnx.proto this importer needs ----------------------------------------------------------------------------

_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 10: np.float16, 11: np.float64}


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    """TensorProto: dims = 1, data_type = 2, float_data = 4, int64_data = 7, name = 8, raw_data = 9."""
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += [_signed(x) for x in _packed_varints(v, wt)]
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            floats += list(np.frombuffer(bytes(v), "<f4")) if wt == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]]
        elif fno == 7:
            ints += [_signed(x) for x in _packed_varints(v, wt)]
    if dtype not in _DTYPES:
        raise OnnxImportError(f"tensor {name!r}: unsupported data_type {dtype}")
    if raw is not None:
        arr = np.frombuffer(raw, np.dtype(_DTYPES[dtype]).newbyteorder("<")).astype(_DTYPES[dtype])
    elif dtype == 1:
        arr = np.asarray(floats, np.float32)
    else:
        arr = np.asarray(ints, _DTYPES[dtype])
    if any(d < 0 for d in dims) or int(np.prod(dims, dtype=np.float64)) != arr.size:
        raise OnnxImportError(f"tensor {name!r}: {arr.size} elements for dims {dims}")
    return name, arr.reshape(dims)


@dataclass
class Node:
    op: str
    name: str
    inputs: List[str]
    outputs: List[str]
    ints: Dict[str, List[int]] = field(default_factory=dict)     # int / ints attributes
    floats: Dict[str, float] = field(default_factory=dict)       # float attributes
    tensors: Dict[str, np.ndarray] = field(default_factory=dict)  # tensor attributes (Constant nodes)


def _attribute(buf: memoryview, node: Node) -> None:
    """AttributeProto: name = 1, f = 2, i = 3, t = 5, ints = 8."""
    name, ints, f, t = "", [], None, None
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            f = struct.unpack("<f", struct.pack("<I", v))[0]
        elif fno == 3:
            ints.append(_signed(v))
        elif fno == 8:
            ints += [_signed(x) for x in _packed_varints(v, wt)]
        elif fno == 5:
            t = _tensor(v)[1]
    if t is not None:
        node.tensors[name] = t
    elif f is not None:
        node.floats[name] = f
    else:
        node.ints[name] = ints


def _node(buf: memoryview) -> Node:
    """NodeProto: input = 1, output = 2, name = 3, op_type = 4, attribute = 5."""
    n = Node("", "", [], [])
    for fno, _, v in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(v).decode())
        elif fno == 2:
            n.outputs.append(bytes(v).decode())
        elif fno == 3:
            n.name = bytes(v).decode()
        elif fno == 4:
            n.op = bytes(v).decode()
        elif fno == 5:
            _attribute(v, n)
    return n


def _value_info(buf: memoryview) -> Tuple[str, List[Optional[int]]]:
    """ValueInfoProto: name = 1, type = 2 { tensor_type = 1 { shape = 2 { dim = 1 { dim_value = 1 } } } }."""
    name, shape = "", []
    for fno, _, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            for f2, _, v2 in _fields(v):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 != 2:
                        continue
                    for f4, _, v4 in _fields(v3):
                        if f4 == 1:
                            dv = [x for f5, _, x in _fields(v4) if f5 == 1]
                            shape.append(_signed(dv[0]) if dv else None)
    return name, shape


@dataclass
class OnnxModel:
    nodes: List[Node]
    initializers: "OrderedDict[str, np.ndarray]"
    inputs: "OrderedDict[str, List[Optional[int]]]"
    outputs: "OrderedDict[str, List[Optional[int]]]"
    opset: int


def _malformed(fn):
    """The file comes from outside: whatever a malformed byte string trips over inside the decoders is reported as OnnxImportError."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        try:
            return fn(*a, **kw)
        except OnnxImportError:
            raise
        except (struct.error, IndexError, KeyError, ValueError, TypeError, OverflowError, UnicodeDecodeError, MemoryError, AttributeError) as e:
            raise OnnxImportError(f"malformed ONNX file ({type(e).__name__}: {e})") from None

    return wrapper


@_malformed
def read_onnx(path: str) -> OnnxModel:
    """ModelProto: graph = 7, opset_import = 8 { version = 2 }; GraphProto: node = 1, initializer = 5, input = 11, output = 12."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    graph, opset = None, 0
    for fno, wt, v in _fields(buf):
        if fno == 7 and wt == 2:
            graph = v
        elif fno == 8 and wt == 2:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    opset = max(opset, v2)
    if graph is None:
        raise OnnxImportError(f"{path}: no GraphProto in the file")
    m = OnnxModel([], OrderedDict(), OrderedDict(), OrderedDict(), opset)
    for fno, wt, v in _fields(graph):
        if wt != 2:
            continue
        if fno == 1:
            m.nodes.append(_node(v))
        elif fno == 5:
            name, arr = _tensor(v)
            m.initializers[name] = arr
        elif fno == 11:
            name, shape = _value_info(v)
            m.inputs[name] = shape
        elif fno == 12:
            name, shape = _value_info(v)
            m.outputs[name] = shape
    for name in list(m.inputs):  # older exporters list the initializers among the graph inputs
        if name in m.initializers:
            del m.inputs[name]
    return m


# ---- architecture detection + mapping onto the VTW1 tensor list -------------------------------------------------------------

def _const(m: OnnxModel, name: str) -> Optional[np.ndarray]:
    if name in m.initializers:
        return m.initializers[name]
    for n in m.nodes:
        if n.op == "Constant" and n.outputs and n.outputs[0] == name and "value" in n.tensors:
            return n.tensors["value"]
    return None


def _linear(m: OnnxModel, mm: Node, consumers: Dict[str, List[Node]]) -> Tuple[np.ndarray, np.ndarray]:
    """torch `nn.Linear` exports as MatMul(x, W^T) + Add(bias) or as Gemm(x, W, b, transB = 1): returns (W [out, in], b [out])."""
    if mm.op == "Gemm":
        w, b = _const(m, mm.inputs[1]), _const(m, mm.inputs[2]) if len(mm.inputs) > 2 else None
        if w is None or b is None:
            raise OnnxImportError(f"Gemm {mm.name!r}: weights are not constants")
        if not mm.ints.get("transB", [0])[0]:
            w = w.T
        return np.ascontiguousarray(w, np.float32), np.asarray(b, np.float32)
    w = _const(m, mm.inputs[1])
    if w is None or w.ndim != 2:
        raise OnnxImportError(f"MatMul {mm.name!r}: the second operand is not a constant 2-D weight")
    for add in consumers.get(mm.outputs[0], []):
        if add.op == "Add":
            other = [i for i in add.inputs if i != mm.outputs[0]]
            b = _const(m, other[0]) if other else None
            if b is not None and b.ndim == 1 and b.shape[0] == w.shape[1]:
                return np.ascontiguousarray(w.T, np.float32), np.asarray(b, np.float32)
    raise OnnxImportError(f"MatMul {mm.name!r}: no bias Add follows it")


def _layernorms(m: OnnxModel) -> List[Tuple[np.ndarray, np.ndarray, float]]:
    out = []
    for n in m.nodes:
        if n.op == "LayerNormalization":
            g, b = _const(m, n.inputs[1]), _const(m, n.inputs[2]) if len(n.inputs) > 2 else None
            if g is None or b is None:
                raise OnnxImportError(f"LayerNormalization {n.name!r}: scale / bias are not constants")
            out.append((np.asarray(g, np.float32), np.asarray(b, np.float32), n.floats.get("epsilon", 1e-5)))
    return out


@_malformed
def detect_and_map(m: OnnxModel) -> Tuple[W.ModelConfig, "OrderedDict[str, np.ndarray]"]:
    """Graph -> (ModelConfig, tensors in `weights.tensor_specs` order).  Raises OnnxImportError on any other topology."""
    ops: Dict[str, int] = {}
    for n in m.nodes:
        ops[n.op] = ops.get(n.op, 0) + 1
    found = ", ".join(f"{k} x{v}" for k, v in sorted(ops.items()))
    want_in = {"template": [1, 3, 128, 128], "search": [1, 3, 256, 256]}
    for name, shape in want_in.items():
        if name not in m.inputs:
            raise OnnxImportError(f"graph input {name!r} missing (inputs: {list(m.inputs)})")
        if [d for d in m.inputs[name]] != shape and None not in m.inputs[name]:
            raise OnnxImportError(f"graph input {name!r} has shape {m.inputs[name]}, expected {shape}")
    if len(m.outputs) != 3:
        raise OnnxImportError(f"expected three output maps, found {list(m.outputs)}")
    consumers: Dict[str, List[Node]] = {}
    for n in m.nodes:
        for i in n.inputs:
            consumers.setdefault(i, []).append(n)

    convs = [n for n in m.nodes if n.op == "Conv"]
    patch = [n for n in convs if n.ints.get("kernel_shape") == [16, 16] and n.ints.get("strides") == [16, 16]]
    if not patch:
        raise OnnxImportError(f"no 16x16 stride-16 patch-embedding Conv in the graph (ops: {found})")
    pw = _const(m, patch[0].inputs[1])
    pb = _const(m, patch[0].inputs[2]) if len(patch[0].inputs) > 2 else None
    if pw is None or pb is None or pw.shape[1:] != (3, 16, 16):
        raise OnnxImportError("patch-embedding Conv: weight / bias are not constants of shape [D, 3, 16, 16] / [D]")
    for p in patch[1:]:  # template and search branches must share the embedding
        w2 = _const(m, p.inputs[1])
        if w2 is None or w2.shape != pw.shape or not np.array_equal(w2, pw):
            raise OnnxImportError("template and search use different patch embeddings: not the one-stream architecture")
    D = int(pw.shape[0])

    depth = ops.get("Softmax", 0)
    lns = _layernorms(m)
    if depth == 0 or len(lns) != 2 * depth + 1:
        raise OnnxImportError(f"expected depth x (2 LayerNormalization + 1 Softmax) + final LayerNormalization; ops: {found}")
    if ops.get("Erf", 0) != depth:
        raise OnnxImportError(f"expected one exact (Erf) GELU per block; ops: {found}")
    for _, _, eps in lns:
        if abs(eps - 1e-6) > 1e-9:
            raise OnnxImportError(f"LayerNormalization epsilon {eps} (the kernels use 1e-6)")

    linears = [(n, *_linear(m, n, consumers)) for n in m.nodes if n.op in ("MatMul", "Gemm") and _const(m, n.inputs[1]) is not None
               and _const(m, n.inputs[1]).ndim == 2]
    if len(linears) != 4 * depth:
        raise OnnxImportError(f"expected 4 weight MatMuls per block ({4 * depth}), found {len(linears)}; ops: {found}")
    hidden = int(linears[2][1].shape[0])

    # position embeddings: the two [1, 64, D] / [1, 256, D] constants added to the patch tokens
    pos = {}
    for name, arr in list(m.initializers.items()) + [(n.outputs[0], n.tensors["value"]) for n in m.nodes if n.op == "Constant" and "value" in n.tensors]:
        a = np.asarray(arr)
        if a.dtype == np.float32 and a.ndim in (2, 3) and a.shape[-1] == D and a.shape[-2] in (64, 256) and a.size == a.shape[-2] * D:
            if any(c.op == "Add" for c in consumers.get(name, [])):
                pos[a.shape[-2]] = a.reshape(a.shape[-2], D)
    if set(pos) != {64, 256}:
        raise OnnxImportError("position embeddings [64, D] (template) and [256, D] (search) not found")

    # heads per block from the Reshape that splits QKV: [B, N, 3, heads, head_dim]
    heads = 0
    for n in m.nodes:
        if n.op == "Reshape":
            shp = _const(m, n.inputs[1])
            if shp is not None and shp.size == 5 and int(shp.reshape(-1)[2]) == 3:
                heads = int(shp.reshape(-1)[3])
                break
    if heads == 0:  # exporter kept the shape dynamic (Concat of scalars): collect the constants that feed it
        for n in m.nodes:
            if n.op == "Concat" and len(n.inputs) == 5:
                vals = [_const(m, i) for i in n.inputs]
                if vals[2] is not None and int(np.asarray(vals[2]).reshape(-1)[0]) == 3 and vals[3] is not None:
                    heads = int(np.asarray(vals[3]).reshape(-1)[0])
                    break
    if heads <= 0 or D % heads or D // heads != 64 and D // heads != 16:
        # the CUDA attention kernels exist for head_dim 64 (tensor cores) and 16 (nano)
        raise OnnxImportError(f"could not determine a supported head count (D = {D}, heads = {heads})")

    head = [n for n in convs if n not in patch]
    if len(head) != 2 or head[0].ints.get("kernel_shape") != [3, 3] or head[1].ints.get("kernel_shape") != [1, 1] or ops.get("Relu", 0) != 1:
        raise OnnxImportError(f"expected a 3x3 Conv + ReLU + 1x1 Conv head; convs: {[n.ints.get('kernel_shape') for n in convs]}")
    h1w, h1b = _const(m, head[0].inputs[1]), _const(m, head[0].inputs[2])
    h2w, h2b = _const(m, head[1].inputs[1]), _const(m, head[1].inputs[2])
    if h1w is None or h2w is None or h1b is None or h2b is None or h2w.shape[0] != W.N_OUT or h1w.shape[1] != D:
        raise OnnxImportError("head convolutions: unexpected weight shapes")
    cfg = W.ModelConfig("imported", D=D, depth=depth, heads=heads, hidden=hidden, head_ch=int(h1w.shape[0]))

    t: "OrderedDict[str, np.ndarray]" = OrderedDict()
    t["patch_w"], t["patch_b"] = pw.reshape(D, W.PATCH_K), pb
    t["pos_z"], t["pos_x"] = pos[64], pos[256]
    for i in range(depth):
        p = f"blk{i}."
        (t[p + "ln1_g"], t[p + "ln1_b"], _), (t[p + "ln2_g"], t[p + "ln2_b"], _) = lns[2 * i], lns[2 * i + 1]
        for k, nm in enumerate(("qkv", "proj", "fc1", "fc2")):
            _, w, b = linears[4 * i + k]
            t[p + nm + "_w"], t[p + nm + "_b"] = w, b
    t["lnf_g"], t["lnf_b"], _ = lns[2 * depth]
    t["head1_w"], t["head1_b"] = h1w, h1b
    t["head2_w"], t["head2_b"] = h2w.reshape(W.N_OUT, -1), h2b
    ordered: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in W.tensor_specs(cfg):
        a = np.ascontiguousarray(t[name], np.float32)
        if a.shape != tuple(shape):
            raise OnnxImportError(f"{name}: shape {a.shape} in the ONNX graph, {tuple(shape)} expected for D={D} hidden={hidden}")
        ordered[name] = a
    return cfg, ordered


def import_onnx(onnx_path: str, vtw_path: str) -> W.ModelConfig:
    """Reads `onnx_path`, writes the VTW1 file `vtw_path`, returns the detected configuration."""
    cfg, tensors = detect_and_map(read_onnx(onnx_path))
    W.save_weights(vtw_path, cfg, tensors)
    return cfg


if __name__ == "__main__":
    import sys

    if len(sys.argv) != 3:
        sys.exit("usage: python -m gstreamer_vit_tracker_b200.onnx_import model.onnx out.vtw")
    c = import_onnx(sys.argv[1], sys.argv[2])
    print(f"D={c.D} depth={c.depth} heads={c.heads} hidden={c.hidden} head_ch={c.head_ch} -> {sys.argv[2]}")
