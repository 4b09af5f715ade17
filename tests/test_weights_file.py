"""Model-file validation (`vt_weights_probe`, the checks `vt_tracker_create` runs before it allocates anything): a file from outside
must never turn into a crash, a huge allocation or an overflow.  CPU only (the entry point does not touch the GPU)."""
import os
import struct

import numpy as np
import pytest

from gstreamer_vit_tracker_b200 import _lib as L, api, weights


@pytest.fixture(scope="module")
def nano_file(built, weight_dir):
    return weights.ensure_weight_file("nano", weight_dir)


def test_probe_reports_the_shape(nano_file):
    m = weights.MODELS["nano"]
    assert api.weights_probe(nano_file) == (m.D, m.depth, m.heads, m.hidden, m.head_ch)


def test_missing_and_foreign_files(tmp_path, built):
    with pytest.raises(api.VtError) as e:
        api.weights_probe(str(tmp_path / "nope.vtw"))
    assert e.value.status == L.VT_ERR_WEIGHTS
    p = tmp_path / "x.vtw"
    p.write_bytes(b"ONNX" + b"\0" * 64)
    with pytest.raises(api.VtError):
        api.weights_probe(str(p))
    with pytest.raises(api.VtError):
        api.weights_probe(str(tmp_path))  # a directory


def test_header_fuzz_never_crashes(nano_file, tmp_path):
    from hypothesis import given, settings, strategies as stt

    blob = open(nano_file, "rb").read()
    p = tmp_path / "fuzz.vtw"
    i32 = stt.one_of(stt.integers(-2**31, 2**31 - 1), stt.sampled_from([0, -1, 1, 32, 64, 192, 2**31 - 1, -2**31, 2**30]))

    @settings(max_examples=200, deadline=None)
    @given(stt.lists(i32, min_size=7, max_size=7), stt.integers(0, 4096))
    def run(hdr, tail):
        p.write_bytes(b"VTW1" + struct.pack("<7i", *hdr) + blob[32:32 + tail])
        try:
            shape = api.weights_probe(str(p))
        except api.VtError as e:
            assert e.status == L.VT_ERR_WEIGHTS
        else:  # accepted: the header must describe exactly this file
            cfg = weights.ModelConfig("x", *shape)
            assert os.path.getsize(p) == 32 + 4 * weights.n_params(cfg)

    run()


def test_truncated_and_padded_files(nano_file, tmp_path):
    blob = open(nano_file, "rb").read()
    for data in (blob[:-4], blob + b"\0\0\0\0", blob[:32], blob[:16]):
        p = tmp_path / "t.vtw"
        p.write_bytes(data)
        with pytest.raises(api.VtError):
            api.weights_probe(str(p))
