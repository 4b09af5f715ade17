"""GPU parity, pixel kernels, through the C ABI (api.py -> libvittrack_b200.so) against the CPU oracle
and the committed fixtures.  Bar: bit-exact."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from gstreamer_vit_tracker_b200 import synth

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def api(built):
    from gstreamer_vit_tracker_b200 import api as _api
    return _api


@pytest.fixture(scope="module")
def oracle(built):
    from oracle import oracle as _o
    return _o


def test_nv12_known_answers(api):
    for k in golden("nv12_kat.json")["pixels"]:
        y, u, v = k["yuv"]
        rgb = api.nv12_full_to_rgb_parallel(np.array([y, y, y, y, u, v, 0, 0], np.uint8), 2, 2)
        assert rgb.reshape(-1, 3).tolist() == [k["rgb"]] * 4, k


def test_nv12_golden_frames_incl_odd_sizes(api):
    for f in golden("nv12_kat.json")["frames"]:
        buf = synth.hash_u8(f["seed"], (f["len"],))
        assert sha(api.nv12_full_to_rgb_parallel(buf, f["w"], f["h"])) == f["sha256"], f


@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080), (3840, 2160), (640, 360), (1000, 562), (333, 77), (48, 2)])
def test_nv12_random_frames_bit_exact(api, oracle, w, h):
    n = w * h + ((h + 1) // 2) * w + 2
    for seed in (11, 12):
        buf = synth.hash_u8(seed * 1000 + w, (n,))
        got = api.nv12_full_to_rgb_parallel(buf, w, h)
        ref = oracle.nv12_to_rgb(buf, w, h, 8)
        assert np.array_equal(got, ref), (w, h, seed, int((got != ref).sum()))


def test_nv12_extreme_values_bit_exact(api, oracle):
    """All 2^24 (Y,U,V) triples: a 4096x4096 frame enumerates Y x (U,V) so every LUT entry and clamp is hit."""
    w = h = 4096
    y = (np.arange(w * h, dtype=np.uint32) % 256).astype(np.uint8)
    uvrow = np.arange(h // 2, dtype=np.uint32)[:, None] * (w // 2) + np.arange(w // 2, dtype=np.uint32)[None, :]
    uv = np.empty((h // 2, w), np.uint8)
    uv[:, 0::2] = (uvrow % 256).astype(np.uint8)
    uv[:, 1::2] = ((uvrow // 256) % 256).astype(np.uint8)
    buf = np.concatenate([y, uv.ravel()])
    got = api.nv12_full_to_rgb_parallel(buf, w, h)
    ref = oracle.nv12_to_rgb(buf, w, h, 8)
    assert np.array_equal(got, ref)


def test_nv12_short_buffer_is_black(api):
    buf = np.full(1280 * 720 * 3 // 2 - 1, 200, np.uint8)
    assert not api.nv12_full_to_rgb_parallel(buf, 1280, 720).any()


def test_nv12_pinned_and_pageable_inputs_agree(api):
    w, h = 1920, 1080
    buf = synth.hash_u8(5, (w * h * 3 // 2,))
    pin = api.PinnedBuffer(buf.size)
    pin.array[:] = buf
    a = api.nv12_full_to_rgb_parallel(buf, w, h)
    b = api.nv12_full_to_rgb_parallel(pin.array, w, h)
    assert np.array_equal(a, b)


def test_nv12_device_batch_matches_single(api, oracle, weight_dir):
    """The batched, device-resident entry (bench `roofline` leg) against the oracle, via torch for device memory."""
    import torch

    from gstreamer_vit_tracker_b200 import weights
    w, h, n = 1920, 1080, 5
    fb = w * h * 3 // 2
    host = synth.hash_u8(99, (n, fb))
    d_in = torch.from_numpy(host).cuda()
    d_out = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
    trk = api.VitTrack.new(weights.ensure_weight_file("nano", weight_dir), w, h)
    torch.cuda.synchronize()
    trk.nv12_to_rgb_device(d_in.data_ptr(), fb, d_out.data_ptr(), w * h * 3, n)
    trk.sync()
    got = d_out.cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], oracle.nv12_to_rgb(host[i], w, h, 8)), i


# ---- overlays -----------------------------------------------------------------------------------
def _apply_nv12(api, d, W, H, op, a):
    {"rect": api.draw_rect_nv12, "cross": api.draw_crosshair_nv12, "text": api.draw_text_nv12, "bg": api.draw_background_nv12,
     "cursor": api.draw_cursor, "sel": api.draw_selection}[op](d, W, H, *a)


def _apply_rgb(api, d, W, H, op, a):
    {"rect": api.draw_rect_rgb, "cross": api.draw_crosshair_rgb, "text": api.draw_text_rgb, "bg": api.draw_background_rgb,
     "cursor": api.draw_cursor_rgb, "sel": api.draw_selection_rgb}[op](d, W, H, *a)


def test_overlay_nv12_golden(api):
    g = golden("overlay_golden.json")
    W, H = g["w"], g["h"]
    for case in g["nv12"]:
        d = synth.hash_u8(g["seed_nv12"], (W * H * 3 // 2,)).copy()
        _apply_nv12(api, d, W, H, case["op"], case["args"])
        assert sha(d) == case["sha256"], case


def test_overlay_rgb_golden(api):
    g = golden("overlay_golden.json")
    W, H = g["w"], g["h"]
    for case in g["rgb"]:
        d = synth.hash_u8(g["seed_rgb"], (W * H * 3,)).copy()
        _apply_rgb(api, d, W, H, case["op"], case["args"])
        assert sha(d) == case["sha256"], case


def test_overlay_text_strict_glyphs(api):
    """RGB text: unknown char ≙ get_glyph panic (src/drawing.rs:99) -> VT_ERR_GLYPH; NV12 text skips and advances."""
    d = np.zeros(160 * 96 * 3, np.uint8)
    with pytest.raises(api.VtError) as e:
        api.draw_text_rgb(d, 160, 96, "A?B", 5, 5, 2, 255)
    assert e.value.status == -6
    assert not d.any()


def test_overlay_1080p_random_commands_vs_oracle(api, oracle):
    """Full-size frames, hash-random geometry incl. boxes hanging over every edge; one command per call and
    a composed list (draw order must be kept where commands overlap)."""
    W, H = 1920, 1080
    hv = synth.hash_u64(777, 4000).astype(np.int64)
    k = 0

    def rnd(lo, hi):
        nonlocal k
        v = lo + int(abs(int(hv[k])) % (hi - lo))
        k += 1
        return v
    base = synth.hash_u8(778, (W * H * 3 // 2,))
    for it in range(12):
        d_gpu, d_ref = base.copy(), base.copy()
        x, y, w, h = rnd(-200, W + 100), rnd(-200, H + 100), rnd(0, 700), rnd(0, 500)
        cmds = [api.overlay_cmd(3, 10, 10, 400, 80, 150), api.overlay_cmd(2, 15, 15, a=2, r=255, text="TRACKING"),
                api.overlay_cmd(2, 15, 40, a=2, r=255, text="FPS: 1234"), api.overlay_cmd(2, 15, 65, a=1, r=200, text="conv:0.1ms trk:0.4ms"),
                api.overlay_cmd(2, 250, 15, a=2, r=255, text="score: 87%"), api.overlay_cmd(0, x, y, w, h, 3, 255),
                api.overlay_cmd(1, x + w // 2, y + h // 2, a=15, r=255), api.overlay_cmd(4, rnd(-50, W + 50), rnd(-50, H + 50)),
                api.overlay_cmd(5, rnd(0, W), rnd(0, H), rnd(0, W), rnd(0, H))]
        trk = api._handle_for(W, H, "nv12")
        trk.overlay(d_gpu, cmds)
        oracle.draw_background_nv12(d_ref, W, H, 10, 10, 400, 80, 150)
        oracle.draw_text_nv12(d_ref, W, H, "TRACKING", 15, 15, 2, 255)
        oracle.draw_text_nv12(d_ref, W, H, "FPS: 1234", 15, 40, 2, 255)
        oracle.draw_text_nv12(d_ref, W, H, "conv:0.1ms trk:0.4ms", 15, 65, 1, 200)
        oracle.draw_text_nv12(d_ref, W, H, "score: 87%", 250, 15, 2, 255)
        oracle.draw_rect_nv12(d_ref, W, H, x, y, w, h, 3, 255)
        oracle.draw_crosshair_nv12(d_ref, W, H, x + w // 2, y + h // 2, 15, 255)
        oracle.draw_cursor_nv12(d_ref, W, H, cmds[7].x, cmds[7].y)
        oracle.draw_selection_nv12(d_ref, W, H, cmds[8].x, cmds[8].y, cmds[8].w, cmds[8].h)
        assert np.array_equal(d_gpu, d_ref), (it, x, y, w, h, int((d_gpu != d_ref).sum()))


def test_overlay_rgb_640x512_vs_oracle(api, oracle):
    W, H = 640, 512
    base = synth.hash_u8(779, (W * H * 3,))
    for (x, y, w, h) in [(100, 100, 64, 48), (-20, -10, 64, 48), (600, 480, 64, 48), (300, 200, 1, 1), (-100, -100, 50, 50), (0, 0, 640, 512)]:
        d_gpu, d_ref = base.copy(), base.copy()
        cmds = [api.overlay_cmd(2, 15, 15, a=2, r=255, text="TRACKING", strict=True), api.overlay_cmd(2, 15, 65, a=1, r=200, text="trk:0.4ms", strict=True),
                api.overlay_cmd(0, x, y, w, h, 3, 0, 255, 0), api.overlay_cmd(1, x + w // 2, y + h // 2, a=15, r=0, g=255, b=0),
                api.overlay_cmd(4, x, y, r=0, g=255, b=0), api.overlay_cmd(5, x, y, x + w, y + h, r=255, g=255, b=0),
                api.overlay_cmd(3, x, y + 100, 50, 20)]
        api._handle_for(W, H, "rgb24").overlay(d_gpu, cmds)
        oracle.draw_text_rgb(d_ref, W, H, "TRACKING", 15, 15, 2, 255)
        oracle.draw_text_rgb(d_ref, W, H, "trk:0.4ms", 15, 65, 1, 200)
        oracle.draw_rect_rgb(d_ref, W, H, x, y, w, h, 3, (0, 255, 0))
        oracle.draw_crosshair_rgb(d_ref, W, H, x + w // 2, y + h // 2, 15, (0, 255, 0))
        oracle.draw_cursor_rgb(d_ref, W, H, x, y)
        oracle.draw_selection_rgb(d_ref, W, H, x, y, x + w, y + h)
        oracle.draw_background_rgb(d_ref, W, H, x, y + 100, 50, 20)
        assert np.array_equal(d_gpu, d_ref), (x, y, w, h, int((d_gpu != d_ref).sum()))
