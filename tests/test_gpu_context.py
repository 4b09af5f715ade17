"""TrackerContext + probe on the GPU against the oracle's TrackerContext driven by the same commands:
states, boxes, scores and the overlaid pixels (HUD strings pinned, since they are timing dependent)."""
import numpy as np
import pytest

from gstreamer_vit_tracker_b200 import synth, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(built):
    from gstreamer_vit_tracker_b200 import api as _api
    return _api


@pytest.fixture(scope="module")
def oracle(built):
    from oracle import oracle as _o
    return _o


def _select(ctx_g, ctx_o, api, moves1, moves2):
    for (c, fast) in moves1:
        ctx_g.handle_command(c, fast)
        ctx_o.handle_command({0: "up", 1: "down", 2: "left", 3: "right"}[c], fast)


class FrameMem:
    """Where the probed frame lives: pageable numpy memory (two-step probe: update, then overlay commands + row copies) or pinned
    memory (one synchronisation per frame: HUD queued with the frame, drawn by the frame's last kernel, mirrored zero-copy)."""

    def __init__(self, api, mode, nbytes):
        self.pin = api.PinnedBuffer(nbytes) if mode != "pageable" else None

    def load(self, fr):
        if self.pin is None:
            return fr.copy()
        self.pin.array[:] = fr
        return self.pin.array


MEM_MODES = ["pageable", "pinned", "pinned_window"]


@pytest.mark.parametrize("mem", MEM_MODES)
def test_context_flow_nv12(api, oracle, weight_dir, mem):
    spec = synth.CONFIGS["cfg1"]
    W, H = spec.width, spec.height
    st = synth.SyntheticStream(spec)
    wpath = weights.ensure_weight_file("nano", weight_dir)
    g = api.TrackerContext.new(wpath, W, H, fmt="nv12", upload_window=(mem == "pinned_window"))
    fm = FrameMem(api, mem, st.frame_bytes())
    otrk = oracle.VitTrack(wpath, threads=8)
    o = oracle.TrackerContext(otrk, W, H)
    U = api.UserCommand
    names = {U.MoveUp: "up", U.MoveDown: "down", U.MoveLeft: "left", U.MoveRight: "right", U.Confirm: "confirm", U.Cancel: "cancel"}

    def cmd(c, fast=False):
        g.handle_command(c, fast)
        o.handle_command(names[c], fast)

    hud = ("FPS: 60", "conv:0.1ms trk:0.5ms")
    n = 0

    def frame():
        nonlocal n
        fr = st.frame(n)
        n += 1
        got = fm.load(fr)
        g.probe(got, hud)
        rgb = oracle.nv12_to_rgb(fr, W, H, 8)
        bb = o.process_frame(rgb)
        # oracle-side overlay, reference order (src/pipeline.rs:125-174)
        ref = fr.copy()
        name = o.state_name()
        oracle.draw_background_nv12(ref, W, H, 10, 10, 400, 80, 150)
        oracle.draw_text_nv12(ref, W, H, name, 15, 15, 2, 255)
        oracle.draw_text_nv12(ref, W, H, hud[0], 15, 40, 2, 255)
        oracle.draw_text_nv12(ref, W, H, hud[1], 15, 65, 1, 200)
        if name == "TRACKING":
            oracle.draw_text_nv12(ref, W, H, "score: %.0f%%" % (o.current_score * 100.0), 250, 15, 2, 255)
        sel = o.selection
        if name.startswith("SELECT"):
            oracle.draw_cursor_nv12(ref, W, H, sel[0], sel[1])
            oracle.draw_selection_nv12(ref, W, H, sel[2], sel[3], sel[0], sel[1], sel[4] == 1)
        box = bb if bb is not None else (o.current_bbox if name == "TRACKING" else None)
        if box is not None:
            oracle.draw_rect_nv12(ref, W, H, box[0], box[1], box[2], box[3], 3, 255)
            oracle.draw_crosshair_nv12(ref, W, H, box[0] + box[2] // 2, box[1] + box[3] // 2, 15, 255)
        assert g.state_name() == name, n
        gb = g.current_bbox
        assert (gb.tuple() if gb else None) == o.current_bbox, n
        assert abs(g.current_score - o.current_score) <= 1e-3, n
        assert np.array_equal(got, ref), (n, name, int((got != ref).sum()))

    frame()                                   # SELECT START, cursor in the centre
    for _ in range(1):
        cmd(U.MoveLeft, True)
    for _ in range(2):
        cmd(U.MoveUp, False)
    cmd(U.Confirm)
    frame()                                   # start point set -> SELECT END
    assert g.state_name() == "SELECT END"
    for _ in range(2):
        cmd(U.MoveRight, True)
    for _ in range(8):
        cmd(U.MoveDown, False)
    frame()                                   # dashed selection visible
    cmd(U.Confirm)
    frame()                                   # init + update on the same frame -> TRACKING
    assert g.state_name() == "TRACKING"
    for _ in range(10):
        frame()
    # loss: the search window leaves the frame (≙ update Err -> Lost, src/tracker_context.rs:134-138), HUD shows LOST, no box
    g.tracker.set_rect((-5000, -5000, 20, 20))
    otrk.rect = (-5000, -5000, 20, 20)
    frame()
    assert g.state_name() == "LOST"
    for _ in range(3):
        frame()
    cmd(U.Cancel)
    frame()
    assert g.state_name() == "SELECT START" and g.current_bbox is None
    # a confirm frame whose init + update fails the gate resets the selection (src/tracker_context.rs:100-109)
    cmd(U.Confirm)
    frame()
    for _ in range(30):
        cmd(U.MoveLeft, True)      # far left: the selection box lies on the background, away from the target
    cmd(U.MoveUp, True)
    cmd(U.Confirm)
    frame()
    assert g.state_name() == o.state_name()


def test_context_lost_and_auto_reset(api, weight_dir):
    """Force a loss by moving the search window outside the frame (≙ update Err -> Lost, 61 frames later auto reset)."""
    spec = synth.StreamSpec("lost", 640, 360, 5, [(280, 150, 80, 60, 2, 1)])
    st = synth.SyntheticStream(spec)
    wpath = weights.ensure_weight_file("nano", weight_dir)
    g = api.TrackerContext.new(wpath, spec.width, spec.height)
    U = api.UserCommand
    g.handle_command(U.Confirm)
    g.process_frame(st.frame(0))
    g.handle_command(U.MoveRight, True)
    g.handle_command(U.MoveDown, True)
    g.handle_command(U.Confirm)
    assert g.process_frame(st.frame(1)) is not None and g.state_name() == "TRACKING"
    g.tracker.set_rect((-5000, -5000, 20, 20))
    assert g.process_frame(st.frame(2)) is None and g.state_name() == "LOST"
    for i in range(61):
        g.process_frame(st.frame(3))
        assert g.state_name() == "LOST"
    g.process_frame(st.frame(3))
    assert g.state_name() == "SELECT START" and g.lost_frames == 61


@pytest.mark.parametrize("mem", MEM_MODES)
def test_context_flow_rgb24(api, oracle, weight_dir, mem):
    """The path main() actually runs (src/pipeline_ir.rs): RGB24 640x512, RGB overlay set."""
    spec = synth.CONFIGS["cfg3"]
    W, H = spec.width, spec.height
    st = synth.SyntheticStream(spec)
    wpath = weights.ensure_weight_file("nano", weight_dir)
    g = api.TrackerContext.new(wpath, W, H, fmt="rgb24", upload_window=(mem == "pinned_window"))
    fm = FrameMem(api, mem, st.frame_bytes())
    otrk = oracle.VitTrack(wpath, threads=8)
    o = oracle.TrackerContext(otrk, W, H)
    U = api.UserCommand
    hud = ("FPS: 60", "trk:0.5ms")
    seq = [(U.Confirm, "confirm", False), None, (U.MoveRight, "right", True), (U.MoveDown, "down", True), (U.Confirm, "confirm", False), None, None, None, None]
    n = 0
    for step in seq:
        if step is not None:
            g.handle_command(step[0], step[2])
            o.handle_command(step[1], step[2])
            continue
        fr = st.frame(n).reshape(-1)
        n += 1
        got = fm.load(fr)
        g.probe(got, hud)
        bb = o.process_frame(fr.reshape(H, W, 3))
        ref = fr.copy()
        name = o.state_name()
        oracle.draw_text_rgb(ref, W, H, name, 15, 15, 2, 255)
        oracle.draw_text_rgb(ref, W, H, hud[0], 15, 40, 2, 255)
        oracle.draw_text_rgb(ref, W, H, hud[1], 15, 65, 1, 200)
        if name == "TRACKING":
            oracle.draw_text_rgb(ref, W, H, "score: %.0f%%" % (o.current_score * 100.0), 200, 15, 2, 255)
        sel = o.selection
        if name.startswith("SELECT"):
            oracle.draw_cursor_rgb(ref, W, H, sel[0], sel[1])
            oracle.draw_selection_rgb(ref, W, H, sel[2], sel[3], sel[0], sel[1], sel[4] == 1)
        box = bb if bb is not None else (o.current_bbox if name == "TRACKING" else None)
        if box is not None:
            oracle.draw_rect_rgb(ref, W, H, box[0], box[1], box[2], box[3], 3, (0, 255, 0))
            oracle.draw_crosshair_rgb(ref, W, H, box[0] + box[2] // 2, box[1] + box[3] // 2, 15, (0, 255, 0))
        assert g.state_name() == name
        assert np.array_equal(got, ref), (n, name, int((got != ref).sum()))
    assert g.state_name() == "TRACKING"


def test_probe_score_text_rounding_on_device(api, oracle, weight_dir):
    """The one-synchronisation probe renders "score: NN%" on the device from the fp32 score (round-half-even of score*100, as Rust's
    {:.0} / printf %.0f): the pinned path must draw the same pixels as the two-step path, which formats the string on the host, over a
    sequence whose scores differ (wild weights) — and through the live (timing dependent) HUD path nothing may crash."""
    spec = synth.CONFIGS["cfg1"]
    W, H = spec.width, spec.height
    st = synth.SyntheticStream(spec)
    wpath = weights.ensure_weight_file("nano", weight_dir, variant="wild")
    a = api.TrackerContext.new(wpath, W, H, fmt="nv12", upload_window=True)
    b = api.TrackerContext.new(wpath, W, H, fmt="nv12")
    U = api.UserCommand
    pin = api.PinnedBuffer(st.frame_bytes())
    hud = ("FPS: 59", "conv:0.0ms trk:0.3ms")
    scores = set()
    for n in range(14):
        if n in (0, 1):
            for c in (a, b):
                if n == 1:
                    c.handle_command(U.MoveRight, True)
                    c.handle_command(U.MoveDown, True)
                c.handle_command(U.Confirm)
        fr = st.frame(n)
        pin.array[:] = fr
        pg = fr.copy()
        a.probe(pin.array, hud)
        b.probe(pg, hud)
        assert a.state_name() == b.state_name() and a.current_score == b.current_score, n
        assert np.array_equal(pin.array, pg), (n, a.state_name(), int((pin.array != pg).sum()))
        scores.add(round(a.current_score * 100))
    assert a.state_name() == "TRACKING" and len(scores) >= 2
    for n in range(14, 18):                      # live HUD strings (timing dependent: no pixel comparison)
        pin.array[:] = st.frame(n)
        a.probe(pin.array)
    assert a.state_name() == "TRACKING" and a.tracker.timing().fps > 0


@pytest.mark.gpu
def test_c_example_matches_the_python_binding(built, weight_dir, tmp_path):
    """examples/track_nv12.c (plain C against include/vt_tracker.h) tracks a synthetic NV12 file and prints the same boxes and scores
    as the ctypes binding: the boundary really is usable from C alone."""
    import os
    import re
    import shutil
    import subprocess

    import numpy as np

    from gstreamer_vit_tracker_b200 import _lib, api, synth, weights

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "track_nv12")
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "track_nv12.c"), "-L", libdir,
                    "-lvittrack_b200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True, capture_output=True)
    spec = synth.CONFIGS["cfg1"]
    st = synth.SyntheticStream(spec)
    w = weights.ensure_weight_file("tiny", weight_dir)
    frames = [np.ascontiguousarray(st.frame(i)).reshape(-1) for i in range(6)]
    path = tmp_path / "frames.nv12"
    with open(path, "wb") as f:
        for fr in frames:
            f.write(fr.tobytes())
    box = st.target_boxes(0)[0]
    out = subprocess.run([exe, w, str(path), str(spec.width), str(spec.height), *map(str, box)], check=True, capture_output=True, text=True).stdout
    got = [(float(m.group(1)), tuple(int(v) for v in m.group(2, 3, 4, 5)))
           for m in re.finditer(r"score ([0-9.]+) box \((-?\d+), (-?\d+), (-?\d+), (-?\d+)\)", out)]
    assert len(got) == 5, out
    trk = api.VitTrack.new(w, spec.width, spec.height, box_overlay=True, upload_window=True)
    pin = api.PinnedBuffer(frames[0].size)
    pin.array[:] = frames[0]
    trk.init(pin.array, api.BBox(*box))
    for i in range(1, 6):
        pin.array[:] = frames[i]
        r = trk.update(pin.array)
        assert tuple(r.bbox) == got[i - 1][1] and abs(r.score - got[i - 1][0]) < 1e-4, (i, r, got[i - 1])


@pytest.mark.gpu
def test_c_stream_group_example(built, weight_dir, tmp_path):
    """examples/stream_group.c: three NV12 files stepped together through vt_tracker_update_streams from plain C; every stream's
    printed boxes / scores equal those of a single-stream handle on the same file."""
    import os
    import re
    import shutil
    import subprocess

    import numpy as np

    from gstreamer_vit_tracker_b200 import _lib, api, synth, weights

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "stream_group")
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "stream_group.c"), "-L", libdir,
                    "-lvittrack_b200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True, capture_output=True)
    w = weights.ensure_weight_file("tiny", weight_dir)
    n, steps = 3, 5
    streams = [synth.SyntheticStream(synth.cfg5_stream(i)) for i in range(n)]
    spec = streams[0].spec
    box = streams[0].target_boxes(0)[0]
    paths = []
    for i, st in enumerate(streams):
        p = tmp_path / f"s{i}.nv12"
        with open(p, "wb") as f:
            for k in range(steps + 1):
                f.write(np.ascontiguousarray(st.frame(k)).tobytes())
        paths.append(str(p))
    out = subprocess.run([exe, w, str(spec.width), str(spec.height), *map(str, box), *paths], check=True, capture_output=True, text=True).stdout
    got = {}
    for m in re.finditer(r"step (\d+) stream (\d+): \S+\s+score ([0-9.]+) box \((-?\d+), (-?\d+), (-?\d+), (-?\d+)\)", out):
        got[(int(m.group(1)), int(m.group(2)))] = (float(m.group(3)), tuple(int(v) for v in m.group(4, 5, 6, 7)))
    assert len(got) == n * steps, out
    for i, st in enumerate(streams):
        trk = api.VitTrack.new(w, spec.width, spec.height, box_overlay=True, upload_window=True)
        pin = api.PinnedBuffer(st.frame_bytes())
        pin.array[:] = np.ascontiguousarray(st.frame(0)).reshape(-1)
        trk.init(pin.array, api.BBox(*box))
        for k in range(1, steps + 1):
            pin.array[:] = np.ascontiguousarray(st.frame(k)).reshape(-1)
            r = trk.update(pin.array)
            assert tuple(r.bbox) == got[(k, i)][1] and abs(r.score - got[(k, i)][0]) < 1e-4, (k, i, r, got[(k, i)])
        trk.close()
        pin.close()
