"""TrackerContext / SelectionState / TimingStats: the C oracle AND the product's host logic (C++ inside
libvittrack_b200.so, no GPU needed) against traces of a pure-Python reading of the reference
(src/tracker_context.rs, src/selection_state.rs, src/timing_stats.rs)."""
import pytest

from conftest import golden
from oracle import oracle

CMD_NAMES = {"up": 0, "down": 1, "left": 2, "right": 3, "confirm": 4, "cancel": 5, "quit": 6}


def _check(step_no, rec, returned, state, score, bbox, sel, lost):
    where = f"step {step_no}: {rec['step']}"
    assert state == rec["state"], where
    assert (list(returned) if returned is not None else None) == rec["returned"], where
    assert abs(score - rec["score"]) < 1e-7, where
    assert (list(bbox) if bbox is not None else None) == rec["bbox"], where
    assert list(sel[:5]) == rec["selection"], where
    assert lost == rec["lost"], where


def test_oracle_state_machine_trace():
    g = golden("state_traces.json")
    ctx = oracle.TrackerContext(None, g["w"], g["h"])
    for i, rec in enumerate(g["trace"]):
        kind, a, b = rec["step"]
        ret = None
        if kind == "cmd":
            ctx.handle_command(a, bool(b))
        elif isinstance(a, str):
            ret = ctx.process_frame(None, scripted=(False, 0.0, (0, 0, 0, 0)), scripted_err=True)
        else:
            ret = ctx.process_frame(None, scripted=(a[0], a[1], tuple(a[2])))
        _check(i, rec, ret, ctx.state_name(), ctx.current_score, ctx.current_bbox, ctx.selection, ctx.lost_frames)


def test_product_state_machine_trace(built):
    from gstreamer_vit_tracker_b200 import api

    g = golden("state_traces.json")
    ctx = api.TrackerContext.scripted(g["w"], g["h"])
    for i, rec in enumerate(g["trace"]):
        kind, a, b = rec["step"]
        ret = None
        if kind == "cmd":
            ctx.handle_command(CMD_NAMES[a], bool(b))
        elif isinstance(a, str):
            ret = ctx.process_scripted(None, err=True)
        else:
            ret = ctx.process_scripted(api.TrackResult(a[0], a[1], tuple(a[2])))
        s = ctx.selection
        sel = (s.cursor_x, s.cursor_y, s.start_x, s.start_y, s.phase)
        bb = ctx.current_bbox
        _check(i, rec, ret.tuple() if ret else None, ctx.state_name(), ctx.current_score, bb.tuple() if bb else None, sel, ctx.lost_frames)
        assert (s.step, s.fast_step) == (10, 50)


def test_selection_bbox_min_side(built):
    """get_bbox(): min side 20 (src/selection_state.rs:39-45)."""
    from gstreamer_vit_tracker_b200 import api

    for rec in golden("state_traces.json")["trace"]:
        cx, cy, sx, sy, _ = rec["selection"]
        assert rec["sel_bbox"] == [min(sx, cx), min(sy, cy), max(abs(sx - cx), 20), max(abs(sy - cy), 20)]
    ctx = api.TrackerContext.scripted(100, 80)
    for _ in range(30):
        ctx.handle_command(api.UserCommand.MoveRight, True)
        ctx.handle_command(api.UserCommand.MoveDown, True)
    assert (ctx.selection.cursor_x, ctx.selection.cursor_y) == (99, 79)  # clamp to w-1 / h-1


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_timing_stats(impl, built):
    g = golden("state_traces.json")["timing"]
    if impl == "oracle":
        make = oracle.TimingStats
    else:
        from gstreamer_vit_tracker_b200 import api

        make = api.TimingStats.new
    for chk in g["checks"]:
        t = make()
        for i in range(chk["n"]):
            t.add_interval(g["intervals"][i])
            t.add_times(g["conv"][i], g["track"][i])
        assert t.fps() == pytest.approx(chk["fps"], rel=1e-12)
        assert t.avg_conv_ms() == pytest.approx(chk["conv_ms"], rel=1e-12)
        assert t.avg_track_ms() == pytest.approx(chk["track_ms"], rel=1e-12)
    t = make()
    t.add_interval(0)
    assert t.fps() == 0.0  # avg == 0 -> 0.0 (src/timing_stats.rs:41-45)


def test_keyboard_map_matches_reference_source(built):
    """vt_command_from_key == the byte -> UserCommand match of src/raw_mode_guard.rs:65-101 for all 256 byte values."""
    import ctypes as C
    import json
    import os

    from gstreamer_vit_tracker_b200 import _lib

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "keymap.json")))
    names = ["MoveUp", "MoveDown", "MoveLeft", "MoveRight", "Confirm", "Cancel", "Quit"]
    L = _lib.lib()
    for byte in range(256):
        cmd, fast = C.c_int32(-1), C.c_int32(-1)
        hit = L.vt_command_from_key(byte, C.byref(cmd), C.byref(fast))
        want = gold.get(str(byte))
        if want is None:
            assert hit == 0, byte
        else:
            assert hit == 1 and names[cmd.value] == want[0] and bool(fast.value) == want[1], (byte, cmd.value, fast.value, want)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_product_vs_oracle_random_walk(built, seed):
    """20,000 random steps (commands, confirmed acquisitions, good / weak / failed / erroring updates, long losses that trigger the
    60-frame auto reset): the product's C++ state machine and the C oracle stay in lock step on every observable
    (returned box, state name, score, bbox, selection, lost counter)."""
    import random

    from gstreamer_vit_tracker_b200 import api

    rng = random.Random(seed)
    W, H = 1920, 1080
    a = oracle.TrackerContext(None, W, H)
    b = api.TrackerContext.scripted(W, H)
    names = list(CMD_NAMES)
    seen = set()
    for step in range(20000):
        u = rng.random()
        if u < 0.45:
            cmd = rng.choice(names[:4] if u < 0.30 else names)
            fast = rng.random() < 0.5
            a.handle_command(cmd, fast)
            b.handle_command(CMD_NAMES[cmd], fast)
            ra = rb = None
        else:
            err = rng.random() < 0.05
            ok = rng.random() < 0.8
            score = rng.choice([0.0, 0.1, 0.2, 0.25, 0.2500001, 0.3, 0.9]) if rng.random() < 0.5 else rng.random()
            if rng.random() < 0.02:  # a long loss
                ok, score = False, 0.0
            bb = (rng.randint(-50, W), rng.randint(-50, H), rng.randint(0, 400), rng.randint(0, 400))
            ra = a.process_frame(None, scripted=(ok, score, bb), scripted_err=err)
            rb = b.process_scripted(api.TrackResult(ok, score, bb), err=err)
            rb = rb.tuple() if rb else None
        s = b.selection
        bbb = b.current_bbox
        assert ra == rb, step
        assert a.state_name() == b.state_name(), step
        assert a.current_score == b.current_score, step
        assert a.current_bbox == (bbb.tuple() if bbb else None), step
        assert tuple(a.selection[:5]) == (s.cursor_x, s.cursor_y, s.start_x, s.start_y, s.phase), step
        assert a.lost_frames == b.lost_frames, step
        seen.add(a.state_name())
    assert seen == {"SELECT START", "SELECT END", "TRACKING", "LOST"}, seen
