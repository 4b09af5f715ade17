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
