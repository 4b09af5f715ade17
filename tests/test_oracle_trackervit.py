"""The oracle's VitTrack against the third-party executable implementation cv2.TrackerVit (OpenCV 4.13):
fixtures in tests/golden/trackervit_{nano,tiny}.json were produced by tools/make_golden.py from an ONNX
export of the same weight file.  Boxes must be equal, scores within 1e-5 (nano: D=64, L=2) / 2e-5 (tiny: D=192, L=12, the
bench model — two fp32 implementations with different summation orders over 12 blocks)."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from gstreamer_vit_tracker_b200 import synth, weights
from oracle import oracle


def _spec(d, name="g"):
    return synth.StreamSpec(name, d["w"], d["h"], d["seed"], [tuple(t) for t in d["targets"]])


SCORE_TOL = {"nano": 1e-5, "tiny": 2e-5}


@pytest.mark.parametrize("model", ["nano", "tiny"])
@pytest.mark.parametrize("variant", ["stable", "wild"])
def test_free_running_sequences_match_cv2(variant, model, weight_dir):
    g = golden(f"trackervit_{model}.json")["models"][variant]
    wpath = weights.ensure_weight_file(model, weight_dir, variant=variant)
    assert hashlib.sha256(open(wpath, "rb").read()).hexdigest() == g["weights_sha256"], "weight generator changed"
    for seq in g["sequences"]:
        spec = _spec(seq["spec"])
        st = synth.SyntheticStream(spec)
        trk = oracle.VitTrack(wpath, threads=8)
        trk.set_threshold(0.2)
        trk.init(oracle.nv12_to_rgb(st.frame(0), spec.width, spec.height, 4), tuple(seq["init_box"]))
        for i, fr in enumerate(seq["frames"]):
            rc, ok, score, bb = trk.update(oracle.nv12_to_rgb(st.frame(i), spec.width, spec.height, 4))
            assert rc == 0 and ok == fr["ok"], (seq["name"], i)
            assert abs(score - fr["score"]) < SCORE_TOL[model], (seq["name"], i, score, fr["score"])
            if ok:
                assert list(bb) == fr["bbox"], (seq["name"], i, bb, fr["bbox"])


@pytest.mark.parametrize("model", ["nano", "tiny"])
@pytest.mark.parametrize("variant", ["stable", "wild"])
def test_single_steps_incl_borders_match_cv2(variant, model, weight_dir):
    g = golden(f"trackervit_{model}.json")["models"][variant]
    wpath = weights.ensure_weight_file(model, weight_dir, variant=variant)
    spec = _spec(g["steps_spec"])
    st = synth.SyntheticStream(spec)
    rgb = [oracle.nv12_to_rgb(st.frame(i), spec.width, spec.height, 2) for i in g["steps_spec"]["frames"]]
    n_err = n_pad = 0
    for s in g["single_steps"]:
        trk = oracle.VitTrack(wpath, threads=8)
        rc = trk.init(rgb[0], tuple(s["box"]))
        if s.get("error"):
            # cv2 raised an ROI assertion: the crop lies outside the frame (App. A.1)
            rc2 = trk.update(rgb[1])[0] if rc == 0 else rc
            assert rc != 0 or rc2 != 0, s
            n_err += 1
            continue
        assert rc == 0, s
        rc, ok, score, bb = trk.update(rgb[1])
        assert rc == 0 and ok == s["ok"], s
        assert abs(score - s["score"]) < SCORE_TOL[model], (s, score)
        if ok:
            assert list(bb) == s["bbox"], (s, bb)
        x, y, w, h = s["box"]
        n_pad += x < 0 or y < 0 or x + w > spec.width or y + h > spec.height
    assert n_pad >= 5  # the fixture does exercise padded crops


def test_custom_norm_switch_matches_cv2_default_std_quirk(weight_dir):
    """SURVEY.md App. A.7 `norm_custom` pinned against the third party: cv2.TrackerVit with its DEFAULT stdvalue divides by a Scalar the
    quaternion way (§8c); configured with that affine map (tests/golden/trackervit_variants.json) the oracle reproduces it."""
    g = golden("trackervit_variants.json")
    wpath = weights.ensure_weight_file(g["model"], weight_dir, variant=g["variant"])
    assert hashlib.sha256(open(wpath, "rb").read()).hexdigest() == g["weights_sha256"]
    for seq in g["sequences"]:
        spec = _spec(seq["spec"])
        st = synth.SyntheticStream(spec)
        trk = oracle.VitTrack(wpath, threads=4)
        trk.set_variant(norm=(g["norm_scale"], g["norm_bias"]))
        trk.init(oracle.nv12_to_rgb(st.frame(0), spec.width, spec.height, 4), tuple(seq["init_box"]))
        for i, fr in enumerate(seq["frames"]):
            rc, ok, score, bb = trk.update(oracle.nv12_to_rgb(st.frame(i), spec.width, spec.height, 4))
            assert rc == 0 and ok == fr["ok"] and abs(score - fr["score"]) < 1e-5, (seq["name"], i, score, fr["score"])
            if ok:
                assert list(bb) == fr["bbox"], (seq["name"], i, bb, fr["bbox"])


def test_variant_switches_change_what_they_say(weight_dir):
    """The other App. A.7 switches (older-OpenCV behaviour, restated from recollection — no executable copy offline): each one changes
    exactly its own stage, and the all-zero variant is the default bit for bit."""
    wpath = weights.ensure_weight_file("nano", weight_dir, variant="wild")
    spec = synth.StreamSpec("v", 640, 360, 77, [(560, 300, 80, 60, 2, 1)])   # search window overhangs the right and bottom edges
    st = synth.SyntheticStream(spec)
    rgb0, rgb1 = (oracle.nv12_to_rgb(st.frame(i), 640, 360, 2) for i in (0, 1))
    box = st.target_boxes(0)[0]

    def run(**kw):
        t = oracle.VitTrack(wpath, threads=4)
        if kw:
            t.set_variant(**kw)
        t.init(rgb0, box)
        r = t.update(rgb1)
        return r, t.last_maps(), t.last_blobs()

    base, explicit_default = run(), run(pad_plus1=False)
    assert base[0] == explicit_default[0] and all(np.array_equal(a, b) for a, b in zip(base[1], explicit_default[1]))
    # window: same raw conf, conf_win = conf * (1 - hann)
    r, maps, _ = run(window=1)
    assert np.array_equal(maps[3], base[1][3]) and not np.array_equal(maps[0], base[1][0])
    h1 = (0.5 * (1 - np.cos(np.float32(2 * np.pi / 17) * np.arange(1, 17, dtype=np.float32)))).astype(np.float32)
    np.testing.assert_allclose(maps[0], maps[3] * (np.float32(1) - np.outer(h1, h1).astype(np.float32).reshape(-1)), rtol=0, atol=1e-7)
    # decode_window: same maps and score, the box is scaled by 4*floor(sqrt(w*h)) instead of ceil(4*sqrt(w*h))
    r, maps, _ = run(decode_window=1)
    assert all(np.array_equal(a, b) for a, b in zip(maps, base[1])) and r[2] == base[0][2]
    c_new, c_old = 4 * int(np.floor(np.sqrt(80 * 60))), int(np.ceil(np.sqrt(80 * 60) * 4))
    assert c_new != c_old

    def decode(c):  # App. A.6 in fp32 with crop-window size c
        f, best = np.float32, int(np.argmax(maps[0]))
        my, mx = divmod(best, 16)
        cx, cy = (f(mx) + maps[2][best]) / f(16), (f(my) + maps[2][256 + best]) / f(16)
        bw, bh = maps[1][best], maps[1][256 + best]
        x0, y0 = box[0] + int((box[2] - c) / 2), box[1] + int((box[3] - c) / 2)
        return tuple(int(np.floor(v)) for v in ((cx - bw / f(2)) * f(c) + f(x0), (cy - bh / f(2)) * f(c) + f(y0), bw * f(c), bh * f(c)))
    assert r[3] == decode(c_new) and base[0][3] == decode(c_old)
    # pad_plus1: the blobs differ only where the crop reaches the right / bottom frame edge: one more zero column / row
    r, _, blobs = run(pad_plus1=True)
    d = blobs[0] != base[2][0]
    assert d.any() and not d[:, :128, :128].any()
