"""The C oracle against fixtures that were NOT produced by it (tools/make_golden.py):
reference formulas evaluated in pure Python, the font parsed from the reference source, a
pure-Python reading of the draw loops, cv2.resize."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from gstreamer_vit_tracker_b200 import synth
from oracle import oracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_nv12_known_answers():
    """SURVEY.md Appendix C + 500 hash-random triples (src/nv12_convert.rs:24-30,124-126,41-43)."""
    kat = golden("nv12_kat.json")["pixels"]
    for k in kat:
        y, u, v = k["yuv"]
        frame = np.array([y, y, y, y, u, v], np.uint8)  # 2x2 NV12
        rgb = oracle.nv12_to_rgb(frame, 2, 2)
        assert rgb.reshape(-1, 3).tolist() == [k["rgb"]] * 4, k


@pytest.mark.parametrize("threads", [1, 3])
def test_nv12_frames_incl_odd_sizes(threads):
    for f in golden("nv12_kat.json")["frames"]:
        buf = synth.hash_u8(f["seed"], (f["len"],))
        assert sha(oracle.nv12_to_rgb(buf, f["w"], f["h"], threads)) == f["sha256"], f


def test_nv12_short_buffer_is_black():
    """len < w*h*3/2 -> all zeros (src/nv12_convert.rs:48-50)."""
    buf = np.full(16 * 8 * 3 // 2 - 1, 200, np.uint8)
    assert not oracle.nv12_to_rgb(buf, 16, 8).any()


def test_glyph_table_matches_reference_source():
    g = golden("glyphs.json")["glyphs"]
    assert len(g) == 40
    for ch, rows in g.items():
        assert oracle.get_glyph(ch).tolist() == rows, ch
    assert oracle.get_glyph("?") is None and oracle.get_glyph("a") is None


_NV12 = {"rect": oracle.draw_rect_nv12, "cross": oracle.draw_crosshair_nv12, "bg": oracle.draw_background_nv12,
         "cursor": oracle.draw_cursor_nv12}
_RGB = {"rect": lambda d, w, h, x, y, rw, rh, t, r, g, b: oracle.draw_rect_rgb(d, w, h, x, y, rw, rh, t, (r, g, b)),
        "cross": lambda d, w, h, cx, cy, s, r, g, b: oracle.draw_crosshair_rgb(d, w, h, cx, cy, s, (r, g, b)),
        "bg": oracle.draw_background_rgb, "cursor": oracle.draw_cursor_rgb}


def test_overlay_nv12_against_python_reading():
    g = golden("overlay_golden.json")
    W, H = g["w"], g["h"]
    for case in g["nv12"]:
        d = synth.hash_u8(g["seed_nv12"], (W * H * 3 // 2,)).copy()
        op, a = case["op"], case["args"]
        if op == "text":
            oracle.draw_text_nv12(d, W, H, *a)
        elif op == "sel":
            oracle.draw_selection_nv12(d, W, H, *a)
        else:
            _NV12[op](d, W, H, *a)
        assert sha(d) == case["sha256"], case


def test_overlay_rgb_against_python_reading():
    g = golden("overlay_golden.json")
    W, H = g["w"], g["h"]
    for case in g["rgb"]:
        d = synth.hash_u8(g["seed_rgb"], (W * H * 3,)).copy()
        op, a = case["op"], case["args"]
        if op == "text":
            oracle.draw_text_rgb(d, W, H, *a)
        elif op == "sel":
            oracle.draw_selection_rgb(d, W, H, *a)
        else:
            _RGB[op](d, W, H, *a)
        assert sha(d) == case["sha256"], case


def test_overlay_hud_composition():
    g = golden("overlay_golden.json")
    W, H = g["w"], g["h"]
    d = synth.hash_u8(g["hud_nv12"]["seed"], (W * H * 3 // 2,)).copy()
    oracle.draw_background_nv12(d, W, H, 10, 10, 400, 80, 150)
    oracle.draw_text_nv12(d, W, H, "TRACKING", 15, 15, 2, 255)
    oracle.draw_text_nv12(d, W, H, "FPS: 60", 15, 40, 2, 255)
    oracle.draw_text_nv12(d, W, H, "conv:0.0ms trk:0.5ms", 15, 65, 1, 200)
    oracle.draw_rect_nv12(d, W, H, 60, 30, 50, 40, 3, 255)
    oracle.draw_crosshair_nv12(d, W, H, 85, 50, 15, 255)
    assert sha(d) == g["hud_nv12"]["sha256"]


def test_resize_bit_exact_with_cv2_golden():
    """OpenCV INTER_LINEAR fixed point incl. up-scales (vertical taps clamp the row index)."""
    for c in golden("resize_golden.json")["cases"]:
        img = synth.hash_u8(c["seed"], (c["src"], c["src"], 3))
        assert sha(oracle.resize_linear(img, c["dst"], c["dst"])) == c["sha256"], c


def test_crop_semantics():
    """App. A.1: c = ceil(sqrt(w*h)*factor), truncating division, zero border, outside -> error."""
    img = synth.hash_u8(3, (90, 120, 3))
    rc, crop = oracle.crop_square(img, (50, 40, 20, 10), 2)
    c = int(np.ceil(np.sqrt(200.0) * 2))
    assert rc == 0 and crop.shape == (c, c, 3)
    x1, y1 = 50 + int((20 - c) / 2), 40 + int((10 - c) / 2)
    assert np.array_equal(crop, img[y1:y1 + c, x1:x1 + c])
    rc, crop = oracle.crop_square(img, (-5, -3, 20, 20), 4)  # partly outside: zero padded
    assert rc == 0 and crop.shape == (80, 80, 3)
    x1, y1 = -5 + int((20 - 80) / 2), -3 + int((20 - 80) / 2)
    assert x1 == -35 and y1 == -33
    assert not crop[:33].any() and not crop[:, :35].any()
    assert np.array_equal(crop[33:, 35:], img[0:47, 0:45])
    assert oracle.crop_square(img, (-500, -500, 40, 40), 4)[0] != 0
    assert oracle.crop_square(img, (2000, 10, 40, 40), 4)[0] != 0


def test_normalize_values():
    hwc = np.zeros((128, 128, 3), np.uint8)
    hwc[0, 0] = (255, 128, 0)
    chw = oracle.normalize_chw(hwc)
    np.testing.assert_allclose(chw[:, 5, 5], [-2.1179, -2.0357, -1.8044], atol=1e-4)  # padded pixels, App. A.4
    np.testing.assert_allclose(chw[:, 0, 0], [(1 - 0.485) / 0.229, (128 / 255 - 0.456) / 0.224, (0 - 0.406) / 0.225], rtol=1e-6)
