"""ONNX -> VTW1 importer: the hand-rolled protobuf reader and the architecture detection, pinned by a
round trip through the stand-in network's own ONNX export (tools/torch_model.py, the file cv2.TrackerVit loads for the oracle's
cross-check).  CPU only."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gstreamer_vit_tracker_b200 import onnx_import as oi, weights  # noqa: E402


def _vi(x):  # varint
    out = b""
    while True:
        b = x & 0x7F
        x >>= 7
        out += bytes([b | (0x80 if x else 0)])
        if not x:
            return out


def _ld(fno, payload):  # length-delimited field
    return _vi(fno << 3 | 2) + _vi(len(payload)) + payload


def test_wire_decoder_on_a_handmade_tensor():
    raw = np.arange(6, dtype="<f4").tobytes()
    t = _vi(1 << 3) + _vi(2) + _vi(1 << 3) + _vi(3) + _vi(2 << 3) + _vi(1) + _ld(8, b"w") + _ld(9, raw)  # dims 2, 3; float; name; raw_data
    name, arr = oi._tensor(memoryview(t))
    assert name == "w" and arr.shape == (2, 3) and arr.dtype == np.float32 and arr[1, 2] == 5.0
    packed = _ld(1, _vi(3) + _vi(2)) + _vi(2 << 3) + _vi(7) + _ld(7, _vi(1) + _vi((1 << 64) - 1) + _vi(300) + _vi(4) + _vi(5) + _vi(6))
    name, arr = oi._tensor(memoryview(packed))  # packed dims, int64_data with a negative value
    assert arr.dtype == np.int64 and arr.tolist() == [[1, -1], [300, 4], [5, 6]]


def test_rejects_garbage_and_foreign_graphs(tmp_path):
    p = tmp_path / "bad.onnx"
    p.write_bytes(b"\xff" * 64)
    with pytest.raises(oi.OnnxImportError):
        oi.read_onnx(str(p))
    # a well-formed model whose graph has other inputs: named in the message, not force-fitted
    vinfo = _ld(1, b"image")
    p.write_bytes(_ld(7, _ld(11, vinfo)))
    with pytest.raises(oi.OnnxImportError, match="template"):
        oi.detect_and_map(oi.read_onnx(str(p)))


@pytest.fixture(scope="module")
def nano_onnx(tmp_path_factory):
    torch_model = pytest.importorskip("torch_model")
    d = tmp_path_factory.mktemp("onnx")
    w = weights.ensure_weight_file("nano", str(d), variant="wild")
    onnx = str(d / "nano.onnx")
    torch_model.export_onnx(w, onnx)
    return w, onnx


def test_round_trip_is_bit_exact(nano_onnx, tmp_path):
    w, onnx = nano_onnx
    m = oi.read_onnx(onnx)
    assert list(m.inputs) == ["template", "search"] and m.inputs["search"] == [1, 3, 256, 256]
    assert [m.outputs[k] for k in m.outputs] == [[1, 1, 16, 16], [1, 2, 16, 16], [1, 2, 16, 16]]
    out = str(tmp_path / "imported.vtw")
    cfg = oi.import_onnx(onnx, out)
    ref = weights.MODELS["nano"]
    assert (cfg.D, cfg.depth, cfg.heads, cfg.hidden, cfg.head_ch) == (ref.D, ref.depth, ref.heads, ref.hidden, ref.head_ch)
    a, b = open(w, "rb").read(), open(out, "rb").read()
    assert a == b  # header and every tensor, byte for byte


def test_architecture_mismatch_is_reported(nano_onnx):
    _, onnx = nano_onnx
    m = oi.read_onnx(onnx)
    m.nodes = [n for n in m.nodes if n.op != "Erf"]  # e.g. a tanh-GELU network
    with pytest.raises(oi.OnnxImportError, match="Erf"):
        oi.detect_and_map(m)
    m = oi.read_onnx(onnx)
    key = next(k for k, v in m.initializers.items() if v.ndim == 3 and v.shape[1] == 64)
    del m.initializers[key]
    with pytest.raises(oi.OnnxImportError, match="position"):
        oi.detect_and_map(m)


def test_reader_never_crashes_on_arbitrary_bytes(tmp_path):
    """Fuzz: any byte string either parses or raises OnnxImportError — nothing else (the file comes from outside)."""
    from hypothesis import given, settings, strategies as stt

    p = tmp_path / "fuzz.onnx"

    @settings(max_examples=300, deadline=None)
    @given(stt.binary(min_size=0, max_size=256))
    def run(blob):
        # half of the cases get a valid outer frame (field 7 = graph) so that the nested decoders are reached
        for data in (blob, _ld(7, blob), _ld(7, _ld(5, blob)), _ld(7, _ld(1, blob)), _ld(7, _ld(11, blob))):
            p.write_bytes(data)
            try:
                m = oi.read_onnx(str(p))
                try:
                    oi.detect_and_map(m)
                except oi.OnnxImportError:
                    pass
            except oi.OnnxImportError:
                pass

    run()
