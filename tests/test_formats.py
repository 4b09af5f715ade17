"""SURVEY.md §8(f) row 1 — the format steps either side of the RGB probe of the path main() really runs
(/root/reference/src/pipeline_ir.rs:27-73): YUY2 -> RGB ingest and the RGB display up-scale.
CPU: the oracle against tests/golden/yuy2_golden.json (pure-Python reading of the reference's BT.601 formulas; cv2.resize hashes).
GPU: the CUDA kernels against the oracle, bit-exact."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gstreamer_vit_tracker_b200 import synth, weights  # noqa: E402
from oracle import oracle as orc  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "yuy2_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hash_bytes(seed, n):
    return (synth.hash_u64(seed, n) & np.uint64(0xFF)).astype(np.uint8)


def test_oracle_yuy2_known_answers_and_frames():
    for k in GOLD["kat"]:
        y, u, v = k["yuv"]
        out = orc.yuy2_to_rgb(np.array([y, u, y, v], np.uint8), 2, 1)
        assert out[0, 0].tolist() == k["rgb"] and out[0, 1].tolist() == k["rgb"], k
    for f in GOLD["frames"]:
        w, h = f["w"], f["h"]
        buf = hash_bytes(f["seed"], orc.yuy2_stride(w) * h)
        for threads in (1, 3):
            assert sha(orc.yuy2_to_rgb(buf, w, h, threads)) == f["sha256"], f
    assert GOLD["max_abs_dev_from_cv2_cvtColor"] <= 19  # only for Y < 16, which OpenCV saturates first; +-1 inside the legal range


def test_oracle_yuy2_short_buffer_is_black():
    buf = hash_bytes(1, 64 * 48 * 2)
    assert not orc.yuy2_to_rgb(buf[:-1], 64, 48).any()


def test_oracle_resize_upscale_matches_cv2_golden():
    for r in GOLD["resize"]:
        src = hash_bytes(r["seed"], r["sw"] * r["sh"] * 3).reshape(r["sh"], r["sw"], 3)
        assert sha(orc.resize_linear(src, r["dw"], r["dh"])) == r["sha256"], r


# ---- GPU ------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def trk(tmp_path_factory):
    from gstreamer_vit_tracker_b200 import api
    wpath = weights.ensure_weight_file("nano", str(tmp_path_factory.mktemp("w")))
    return api.VitTrack.new(wpath, 640, 512, fmt="rgb24")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(640, 512), (1920, 1080), (64, 48), (40, 6), (33, 5), (2, 2), (1, 3), (72, 10), (1000, 3)])
def test_gpu_yuy2_bit_exact(trk, w, h):
    buf = hash_bytes(7000 + w + h, orc.yuy2_stride(w) * h)
    got = trk.yuy2_to_rgb(buf, w, h)
    assert np.array_equal(got, orc.yuy2_to_rgb(buf, w, h, 4)), (w, h)


@pytest.mark.gpu
def test_gpu_yuy2_golden_and_short_buffer(trk):
    for f in GOLD["frames"]:
        buf = hash_bytes(f["seed"], orc.yuy2_stride(f["w"]) * f["h"])
        assert sha(trk.yuy2_to_rgb(buf, f["w"], f["h"])) == f["sha256"], f
    assert not trk.yuy2_to_rgb(hash_bytes(3, 640 * 512 * 2 - 4), 640, 512).any()


@pytest.mark.gpu
def test_gpu_yuy2_device_batch(trk):
    import torch
    w, h, n = 640, 512, 5
    fb, ob = w * h * 2, w * h * 3
    buf = hash_bytes(77, fb * n)
    d_in = torch.from_numpy(buf).cuda()
    d_out = torch.zeros(ob * n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    trk.yuy2_to_rgb_device(d_in.data_ptr(), fb, d_out.data_ptr(), ob, w, h, n)
    trk.sync()
    out = d_out.cpu().numpy().reshape(n, h, w, 3)
    for i in range(n):
        assert np.array_equal(out[i], orc.yuy2_to_rgb(buf[i * fb:(i + 1) * fb], w, h, 4)), i


@pytest.mark.gpu
def test_gpu_resize_matches_cv2_golden_and_oracle(trk):
    for r in GOLD["resize"]:
        src = hash_bytes(r["seed"], r["sw"] * r["sh"] * 3).reshape(r["sh"], r["sw"], 3)
        got = trk.resize_rgb(src, r["dw"], r["dh"])
        assert sha(got) == r["sha256"], r
    src = hash_bytes(5, 123 * 77 * 3).reshape(77, 123, 3)
    for dw, dh in [(246, 154), (300, 200), (61, 38), (123, 77), (1, 1)]:
        assert np.array_equal(trk.resize_rgb(src, dw, dh), orc.resize_linear(src, dw, dh)), (dw, dh)


@pytest.mark.gpu
@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 512, 1280, 1024), (320, 200, 1280, 720), (1280, 720, 640, 352), (100, 60, 16, 9), (64, 48, 160, 96)])
def test_gpu_resize_device_batch(trk, sw, sh, dw, dh):
    """vt_resize_rgb_device_batch (precomputed taps, one launch for n frames; the 16-byte store path needs dw*3 % 16 == 0, other
    widths take the per-pixel kernel) against the oracle, up- and down-scales."""
    import torch
    n = 3
    src = hash_bytes(900 + sw + dw, sw * sh * 3 * n).reshape(n, sh, sw, 3)
    d_in = torch.from_numpy(src.copy()).cuda()
    d_out = torch.zeros((n, dh, dw, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    trk.resize_rgb_device_batch(d_in.data_ptr(), sw * sh * 3, sw, sh, d_out.data_ptr(), dw * dh * 3, dw, dh, n)
    trk.sync()
    out = d_out.cpu().numpy()
    for i in range(n):
        assert np.array_equal(out[i], orc.resize_linear(np.ascontiguousarray(src[i]), dw, dh)), (i, sw, sh, dw, dh)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720), (3840, 2160), (640, 512), (528, 4), (512, 6), (16, 4)])
def test_gpu_nv12_wide_and_narrow_kernels_agree(w, h, tmp_path_factory):
    """The 16-byte / four-row NV12->RGB kernel (W % 16 == 0, H % 4 == 0) and the 8-byte / two-row one (H % 4 != 0) against the oracle,
    device-resident batch of 3 frames incl. a partial last 512-px segment."""
    import torch
    from gstreamer_vit_tracker_b200 import api
    wpath = weights.ensure_weight_file("nano", str(tmp_path_factory.mktemp("w")))
    t = api.VitTrack.new(wpath, w, h, fmt="nv12")
    n, fb, ob = 3, w * h * 3 // 2, w * h * 3
    buf = hash_bytes(4000 + w + h, fb * n)
    d_in = torch.from_numpy(buf).cuda()
    d_out = torch.zeros(ob * n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t.nv12_to_rgb_device(d_in.data_ptr(), fb, d_out.data_ptr(), ob, n)
    t.sync()
    out = d_out.cpu().numpy().reshape(n, h, w, 3)
    for i in range(n):
        assert np.array_equal(out[i], orc.nv12_to_rgb(buf[i * fb:(i + 1) * fb], w, h, 4)), (i, w, h)


@pytest.mark.gpu
def test_gpu_ir_pipeline_chain(trk):
    """YUY2 640x512 -> RGB -> (probe: track + overlay on RGB24, covered elsewhere) -> 1280x1024 display frame, all on the GPU path,
    equals the oracle chain byte for byte."""
    w, h = 640, 512
    st = synth.SyntheticStream(synth.CONFIGS["cfg3"])
    rgb_src = np.asarray(st.frame(0)).reshape(h, w, 3)
    # synthesise a YUY2 frame from the RGB test frame (BT.601 limited range forward transform, 2x1 chroma mean)
    r, g, b = [rgb_src[..., i].astype(np.int32) for i in range(3)]
    y = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16
    u = ((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128
    v = ((112 * r - 94 * g - 18 * b + 128) >> 8) + 128
    yuy2 = np.empty((h, w // 2, 4), np.uint8)
    yuy2[..., 0], yuy2[..., 2] = y[:, 0::2], y[:, 1::2]
    yuy2[..., 1] = (u[:, 0::2] + u[:, 1::2] + 1) >> 1
    yuy2[..., 3] = (v[:, 0::2] + v[:, 1::2] + 1) >> 1
    rgb = trk.yuy2_to_rgb(yuy2, w, h)
    assert np.array_equal(rgb, orc.yuy2_to_rgb(yuy2, w, h, 4))
    assert np.abs(rgb.astype(int) - rgb_src.astype(int)).mean() < 4.0  # the round trip through 4:2:2 stays close to the source
    disp = trk.resize_rgb(rgb, 1280, 1024)
    assert np.array_equal(disp, orc.resize_linear(rgb, 1280, 1024))
