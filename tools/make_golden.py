#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ (run in the authoring container only).

Sources of truth, none of which is the C oracle:
  * nv12_kat.json        — pure-Python evaluation of the reference formulas
                           (/root/reference/src/nv12_convert.rs:24-30,124-126,41-43) incl. SURVEY.md Appendix C;
  * glyphs.json          — the 5x7 font parsed out of /root/reference/src/drawing.rs:53-94 and
                           src/nv12_convert.rs:255-296 (asserted identical);
  * overlay_golden.json  — sha256 of frames drawn by a line-by-line pure-Python reading of the reference's
                           draw_* loops (src/nv12_convert.rs:172-343, src/drawing.rs:5-50, src/drawing_rgb.rs:30-128);
  * state_traces.json    — traces of a pure-Python reading of TrackerContext / SelectionState / TimingStats
                           (src/tracker_context.rs, src/selection_state.rs, src/timing_stats.rs);
  * resize_golden.json   — sha256 of cv2.resize(INTER_LINEAR) outputs (OpenCV 4.13) on hash-generated inputs;
  * yuy2_golden.json     — YUY2 -> RGB known answers / frame hashes from a pure-Python evaluation of the reference's BT.601 integer
                           formulas on packed 4:2:2, and cv2.resize(INTER_LINEAR) up-scale hashes;
  * keymap.json          — the keyboard byte -> UserCommand table parsed out of /root/reference/src/raw_mode_guard.rs:65-101;
  * trackervit_nano.json, trackervit_tiny.json — boxes and scores produced by the third-party cv2.TrackerVit (OpenCV 4.13, DNN CPU
                           backend) running an ONNX export of the same weight file; `tiny` is the bench model.
The tests compare the C oracle (and, on the GPU, the CUDA path) with these files.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gstreamer_vit_tracker_b200 import synth, weights  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/src"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------------------------
def yuv_to_rgb(y, u, v):
    def clamp(x):
        return 0 if x < 0 else (255 if x > 255 else x)
    yv = 298 * (y - 16)
    return [clamp((yv + 409 * (v - 128) + 128) >> 8), clamp((yv - 100 * (u - 128) - 208 * (v - 128) + 128) >> 8),
            clamp((yv + 516 * (u - 128) + 128) >> 8)]


def gen_nv12_kat():
    appendix_c = [((16, 128, 128), (0, 0, 0)), ((235, 128, 128), (255, 255, 255)), ((128, 128, 128), (130, 130, 130)),
                  ((17, 128, 128), (1, 1, 1)), ((81, 90, 240), (255, 0, 0)), ((145, 54, 34), (0, 255, 1)), ((41, 240, 110), (0, 0, 255)),
                  ((126, 100, 200), (243, 80, 72)), ((0, 0, 0), (0, 135, 0)), ((255, 255, 255), (255, 125, 255)),
                  ((255, 0, 0), (74, 255, 20)), ((0, 255, 255), (184, 0, 237))]
    for (yuv, rgb) in appendix_c:
        assert tuple(yuv_to_rgb(*yuv)) == rgb, (yuv, rgb, yuv_to_rgb(*yuv))
    rnd = synth.hash_u8(4242, (500, 3))
    kat = [{"yuv": list(map(int, yuv)), "rgb": list(rgb)} for yuv, rgb in appendix_c]
    kat += [{"yuv": [int(a), int(b), int(c)], "rgb": yuv_to_rgb(int(a), int(b), int(c))} for a, b, c in rnd]
    # whole small frames incl. odd sizes: frame bytes from the hash, expected output from the python formulas
    frames = []
    for (w, h, seed) in [(16, 8, 1), (32, 6, 2), (7, 5, 3), (33, 9, 4), (18, 4, 5), (1, 1, 6), (2, 2, 7), (48, 2, 8)]:
        n_uv_rows = (h + 1) // 2
        buf = synth.hash_u8(seed, (w * h + n_uv_rows * w + 2,))
        out = np.zeros((h, w, 3), np.uint8)
        for r in range(h):
            for c in range(w):
                uvi = w * h + (r // 2) * w + (c & ~1)
                out[r, c] = yuv_to_rgb(int(buf[r * w + c]), int(buf[uvi]), int(buf[uvi + 1]))
        frames.append({"w": w, "h": h, "seed": seed, "len": int(buf.size), "sha256": sha(out)})
    json.dump({"source": "src/nv12_convert.rs:24-30,124-126,41-43; SURVEY.md Appendix C", "pixels": kat, "frames": frames},
              open(os.path.join(GOLD, "nv12_kat.json"), "w"))


# ---------------------------------------------------------------------------------------------
def parse_font(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r'\("(.)",\s*\[([^\]]*)\]\)', txt):
        rows = [int(t.strip().replace("0b", ""), 2) for t in m.group(2).split(",") if t.strip()]
        assert len(rows) == 7
        out[m.group(1)] = rows
    return out


def gen_glyphs():
    a, b = parse_font(os.path.join(REF, "drawing.rs")), parse_font(os.path.join(REF, "nv12_convert.rs"))
    assert a == b and len(a) == 40
    json.dump({"source": "src/drawing.rs:53-94 == src/nv12_convert.rs:255-296", "glyphs": a}, open(os.path.join(GOLD, "glyphs.json"), "w"))
    return a


# ---------------------------------------------------------------------------------------------
# Pure-Python reading of the reference draw loops.  usize arithmetic: Python ints + explicit wrap.
U64 = 1 << 64


def as_usize(i32):
    return i32 % U64


def wrap_i32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def sat_sub(a, b):
    return a - b if a > b else 0


def py_draw_rect_nv12(d, width, height, x, y, w, h, thickness, brightness):
    x1, y1 = max(x, 0), max(y, 0)
    x2 = min(as_usize(wrap_i32(x + w)), sat_sub(width, 1))
    y2 = min(as_usize(wrap_i32(y + h)), sat_sub(height, 1))
    for t in range(thickness):
        if y1 + t < height:
            for px in range(x1, x2 + 1):
                d[(y1 + t) * width + px] = brightness
        if y2 >= t and y2 - t < height:
            for px in range(x1, x2 + 1):
                d[(y2 - t) * width + px] = brightness
    for py in range(y1, y2 + 1):
        for t in range(thickness):
            if x1 + t < width:
                d[py * width + x1 + t] = brightness
            if x2 >= t and x2 - t < width:
                d[py * width + x2 - t] = brightness


def py_draw_crosshair_nv12(d, width, height, cx, cy, size, brightness):
    cx, cy = max(cx, 0), max(cy, 0)
    if cy < height:
        for x in range(sat_sub(cx, size), min(cx + size, width - 1) + 1):
            d[cy * width + x] = brightness
    if cx < width:
        for y in range(sat_sub(cy, size), min(cy + size, height - 1) + 1):
            d[y * width + cx] = brightness


def py_draw_text_nv12(font, d, width, height, text, x, y, scale, brightness):
    cursor_x = x
    for ch in text:
        g = font.get(ch)
        if g is not None:
            for row, bits in enumerate(g):
                for col in range(5):
                    if (bits >> (4 - col)) & 1:
                        for dy in range(scale):
                            for dx in range(scale):
                                px, py = cursor_x + col * scale + dx, y + row * scale + dy
                                if px < width and py < height:
                                    d[py * width + px] = brightness
        cursor_x += 6 * scale


def py_draw_background_nv12(d, width, height, x, y, w, h, darkness):
    factor = 255 - darkness
    for py in range(y, min(y + h, height)):
        for px in range(x, min(x + w, width)):
            d[py * width + px] = (int(d[py * width + px]) * factor) // 255


def py_draw_cursor(d, w, h, x, y):
    x, y = min(max(x, 0), w - 1), min(max(y, 0), h - 1)
    for px in range(sat_sub(x, 25), min(x + 25, w - 1) + 1):
        if not (sat_sub(x, 5) <= px <= x + 5):
            d[y * w + px] = 255
    for py in range(sat_sub(y, 25), min(y + 25, h - 1) + 1):
        if not (sat_sub(y, 5) <= py <= y + 5):
            d[py * w + x] = 255


def py_draw_selection(d, w, h, sx, sy, cx, cy):
    x1, y1 = max(min(sx, cx), 0), max(min(sy, cy), 0)
    x2, y2 = min(as_usize(max(sx, cx)), w - 1), min(as_usize(max(sy, cy)), h - 1)
    for x in range(x1, x2 + 1):
        if (x // 6) % 2 == 0:
            d[y1 * w + x] = 255
            d[y2 * w + x] = 255
    for y in range(y1, y2 + 1):
        if (y // 6) % 2 == 0:
            d[y * w + x1] = 255
            d[y * w + x2] = 255


def set_px(d, w, h, x, y, r, g, b):
    if x < 0 or y < 0 or x >= w or y >= h:
        return
    off = (y * w + x) * 3
    if off + 2 < d.size:
        d[off], d[off + 1], d[off + 2] = r, g, b


def py_draw_rect_rgb(d, w, h, x, y, rw, rh, thickness, r, g, b):
    for t in range(thickness):
        for i in range(rw):
            set_px(d, w, h, x + i, y + t, r, g, b)
            set_px(d, w, h, x + i, y + rh - 1 - t, r, g, b)
        for i in range(rh):
            set_px(d, w, h, x + t, y + i, r, g, b)
            set_px(d, w, h, x + rw - 1 - t, y + i, r, g, b)


def py_draw_crosshair_rgb(d, w, h, cx, cy, size, r, g, b):
    for i in range(-size, size + 1):
        set_px(d, w, h, cx + i, cy, r, g, b)
        set_px(d, w, h, cx, cy + i, r, g, b)


def py_draw_cursor_rgb(d, w, h, cx, cy):
    for i in range(5, 26):
        set_px(d, w, h, cx + i, cy, 0, 255, 0)
        set_px(d, w, h, cx - i, cy, 0, 255, 0)
        set_px(d, w, h, cx, cy + i, 0, 255, 0)
        set_px(d, w, h, cx, cy - i, 0, 255, 0)


def py_draw_text_rgb(font, d, w, h, text, x, y, scale, luma):
    cx = x
    for ch in text:
        g = font[ch]  # the reference panics on unknown chars
        for gy, bits in enumerate(g):
            for gx in range(5):
                if (bits >> (4 - gx)) & 1:
                    for sy in range(scale):
                        for sx in range(scale):
                            set_px(d, w, h, cx + gx * scale + sx, y + gy * scale + sy, luma, luma, luma)
        cx += 6 * scale


def py_draw_selection_rgb(d, w, h, sx, sy, cx, cy):
    x1, y1 = max(min(sx, cx), 0), max(min(sy, cy), 0)
    x2, y2 = min(max(sx, cx), w - 1), min(max(sy, cy), h - 1)
    for x in range(x1, x2 + 1):
        if (x // 6) % 2 == 0:
            set_px(d, w, h, x, y1, 255, 255, 0)
            set_px(d, w, h, x, y2, 255, 255, 0)
    for y in range(y1, y2 + 1):
        if (y // 6) % 2 == 0:
            set_px(d, w, h, x1, y, 255, 255, 0)
            set_px(d, w, h, x2, y, 255, 255, 0)


def py_draw_background_rgb(d, w, h, x, y, bw, bh):
    xs, xe = max(x, 0), min(as_usize(wrap_i32(x + bw)), w)
    ys, ye = max(y, 0), min(as_usize(wrap_i32(y + bh)), h)
    for row in range(ys, ye):
        off = (row * w + xs) * 3
        d[off:off + (xe - xs) * 3] = 30


OVERLAY_CASES_NV12 = [
    # (name, args) on a 160x96 frame
    ("rect", (20, 10, 60, 40, 3, 255)), ("rect", (-10, -5, 50, 30, 3, 200)), ("rect", (130, 70, 60, 60, 3, 255)),
    ("rect", (-40, -40, 20, 20, 3, 255)), ("rect", (200, 30, 10, 10, 2, 99)), ("rect", (30, 200, 10, 10, 3, 77)),
    ("rect", (50, 50, 0, 0, 3, 255)), ("rect", (0, 0, 159, 95, 1, 128)), ("rect", (10, 90, 30, 30, 5, 250)),
    ("cross", (80, 48, 15, 255)), ("cross", (3, 2, 15, 255)), ("cross", (158, 94, 15, 9)), ("cross", (-7, 40, 15, 255)), ("cross", (300, 40, 15, 255)),
    ("text", ("TRACKING", 15, 15, 2, 255)), ("text", ("FPS: 60", 5, 40, 2, 255)), ("text", ("conv:1.5ms trk:0.7ms", 3, 65, 1, 200)),
    ("text", ("score: 93%", 100, 80, 2, 255)), ("text", ("A?B", 10, 10, 3, 255)), ("text", ("SELECT START", 100, 0, 2, 255)),
    ("bg", (10, 10, 100, 50, 150)), ("bg", (100, 60, 400, 80, 150)), ("bg", (0, 0, 160, 96, 255)), ("bg", (5, 5, 10, 10, 0)),
    ("cursor", (80, 48)), ("cursor", (2, 3)), ("cursor", (159, 95)), ("cursor", (-20, 500)),
    ("sel", (20, 20, 100, 70)), ("sel", (100, 70, 20, 20)), ("sel", (0, 0, 159, 95)), ("sel", (50, 50, 50, 50)),
]
OVERLAY_CASES_RGB = [
    ("rect", (20, 10, 60, 40, 3, 0, 255, 0)), ("rect", (-10, -5, 50, 30, 3, 0, 255, 0)), ("rect", (130, 70, 60, 60, 3, 9, 8, 7)),
    ("rect", (-40, -40, 20, 20, 3, 0, 255, 0)), ("rect", (50, 50, 0, 0, 3, 1, 2, 3)), ("rect", (40, 40, 2, 2, 3, 1, 2, 3)),
    ("cross", (80, 48, 15, 0, 255, 0)), ("cross", (3, 2, 15, 0, 255, 0)), ("cross", (158, 94, 15, 5, 6, 7)), ("cross", (-7, 40, 15, 0, 255, 0)),
    ("text", ("TRACKING", 15, 15, 2, 255)), ("text", ("trk:0.7ms", 15, 65, 1, 200)), ("text", ("score: 93%", 100, 80, 2, 255)), ("text", ("LOST", -8, -3, 3, 255)),
    ("bg", (10, 10, 100, 50)), ("bg", (100, 60, 400, 80)), ("bg", (-5, -5, 20, 20)),
    ("cursor", (80, 48)), ("cursor", (2, 3)), ("cursor", (159, 95)), ("cursor", (-10, 40)),
    ("sel", (20, 20, 100, 70)), ("sel", (100, 70, 20, 20)), ("sel", (0, 0, 159, 95)),
]


def gen_overlay(font):
    W, H = 160, 96
    out = {"w": W, "h": H, "seed_nv12": 77, "seed_rgb": 78, "nv12": [], "rgb": []}
    for name, args in OVERLAY_CASES_NV12:
        d = synth.hash_u8(77, (W * H * 3 // 2,)).copy()
        if name == "rect": py_draw_rect_nv12(d, W, H, *args)
        elif name == "cross": py_draw_crosshair_nv12(d, W, H, *args)
        elif name == "text": py_draw_text_nv12(font, d, W, H, *args)
        elif name == "bg": py_draw_background_nv12(d, W, H, *args)
        elif name == "cursor": py_draw_cursor(d, W, H, *args)
        elif name == "sel": py_draw_selection(d, W, H, *args)
        out["nv12"].append({"op": name, "args": list(args), "sha256": sha(d)})
    for name, args in OVERLAY_CASES_RGB:
        d = synth.hash_u8(78, (W * H * 3,)).copy()
        if name == "rect": py_draw_rect_rgb(d, W, H, *args)
        elif name == "cross": py_draw_crosshair_rgb(d, W, H, *args)
        elif name == "text": py_draw_text_rgb(font, d, W, H, *args)
        elif name == "bg": py_draw_background_rgb(d, W, H, *args)
        elif name == "cursor": py_draw_cursor_rgb(d, W, H, *args)
        elif name == "sel": py_draw_selection_rgb(d, W, H, *args)
        out["rgb"].append({"op": name, "args": list(args), "sha256": sha(d)})
    # the composed HUD of one NV12 probe frame (src/pipeline.rs:125-168) in reference order
    d = synth.hash_u8(79, (W * H * 3 // 2,)).copy()
    py_draw_background_nv12(d, W, H, 10, 10, 400, 80, 150)
    py_draw_text_nv12(font, d, W, H, "TRACKING", 15, 15, 2, 255)
    py_draw_text_nv12(font, d, W, H, "FPS: 60", 15, 40, 2, 255)
    py_draw_text_nv12(font, d, W, H, "conv:0.0ms trk:0.5ms", 15, 65, 1, 200)
    py_draw_rect_nv12(d, W, H, 60, 30, 50, 40, 3, 255)
    py_draw_crosshair_nv12(d, W, H, 85, 50, 15, 255)
    out["hud_nv12"] = {"seed": 79, "sha256": sha(d)}
    json.dump(out, open(os.path.join(GOLD, "overlay_golden.json"), "w"))


# ---------------------------------------------------------------------------------------------
class PySelection:  # src/selection_state.rs
    def __init__(self, w, h):
        self.cursor_x = self.start_x = w // 2
        self.cursor_y = self.start_y = h // 2
        self.phase, self.step, self.fast_step = 0, 10, 50

    def move(self, dx, dy, fast, w, h):
        s = self.fast_step if fast else self.step
        self.cursor_x = min(max(self.cursor_x + dx * s, 0), w - 1)
        self.cursor_y = min(max(self.cursor_y + dy * s, 0), h - 1)

    def bbox(self):
        return [min(self.start_x, self.cursor_x), min(self.start_y, self.cursor_y),
                max(abs(self.start_x - self.cursor_x), 20), max(abs(self.start_y - self.cursor_y), 20)]


class PyContext:  # src/tracker_context.rs
    def __init__(self, w, h):
        self.w, self.h = w, h
        self.state, self.lost = "Selecting", 0
        self.sel = PySelection(w, h)
        self.bbox, self.score, self.pending = None, 0.0, False

    def cmd(self, c, fast):
        if c == "up": self.sel.move(0, -1, fast, self.w, self.h)
        elif c == "down": self.sel.move(0, 1, fast, self.w, self.h)
        elif c == "left": self.sel.move(-1, 0, fast, self.w, self.h)
        elif c == "right": self.sel.move(1, 0, fast, self.w, self.h)
        elif c == "confirm": self.pending = True
        elif c == "cancel":
            self.state, self.sel, self.bbox = "Selecting", PySelection(self.w, self.h), None

    def frame(self, upd):  # upd = None (Err) or (success, score, bbox)
        if self.state == "Selecting":
            if self.pending:
                self.pending = False
                if self.sel.phase == 0:
                    self.sel.start_x, self.sel.start_y, self.sel.phase = self.sel.cursor_x, self.sel.cursor_y, 1
                else:
                    if upd is not None and upd[0] and np.float32(upd[1]) > np.float32(0.25):
                        self.bbox, self.score, self.state = list(upd[2]), upd[1], "Tracking"
                        return self.bbox
                    self.sel = PySelection(self.w, self.h)
            return None
        if self.state == "Tracking":
            self.pending = False
            if upd is not None:
                if upd[0] and np.float32(upd[1]) > np.float32(0.25):
                    self.bbox, self.score = list(upd[2]), upd[1]
                    return self.bbox
                self.state, self.lost, self.score = "Lost", 0, 0.0
                return None
            self.state, self.lost = "Lost", 0
            return None
        self.pending = False
        if self.lost > 60:
            self.state, self.sel, self.bbox = "Selecting", PySelection(self.w, self.h), None
        else:
            self.lost += 1
        return None

    def name(self):
        if self.state == "Selecting":
            return "SELECT START" if self.sel.phase == 0 else "SELECT END"
        return "TRACKING" if self.state == "Tracking" else "LOST"


def gen_state_traces():
    W, H = 1920, 1080
    script = []
    script += [("cmd", "left", True)] * 3 + [("cmd", "up", False)] * 4 + [("cmd", "confirm", False), ("frame", "none")]
    script += [("cmd", "right", True)] * 2 + [("cmd", "down", True)] + [("cmd", "confirm", False), ("frame", (True, 0.9, (10, 20, 30, 40)))]
    script += [("frame", (True, 0.8, (11, 21, 31, 41))), ("cmd", "confirm", False), ("frame", (True, 0.26, (12, 22, 32, 42)))]
    script += [("frame", (True, 0.25, (13, 23, 33, 43)))]  # not > 0.25 -> Lost
    script += [("frame", "none")] * 63  # Lost frames 1..61, auto reset, then selecting
    script += [("cmd", "confirm", False), ("frame", "none"), ("cmd", "confirm", False), ("frame", (True, 0.2, (1, 2, 3, 4)))]  # low score -> reset
    script += [("cmd", "confirm", False), ("frame", "none"), ("cmd", "confirm", False), ("frame", "err")]  # Err -> reset
    script += [("cmd", "left", True)] * 30 + [("cmd", "up", True)] * 30  # clamp at 0
    script += [("cmd", "confirm", False), ("frame", "none"), ("cmd", "right", False), ("cmd", "confirm", False),
               ("frame", (True, 0.5, (5, 6, 20, 20))), ("frame", "err"), ("frame", "none"), ("cmd", "cancel", False), ("frame", "none")]
    script += [("cmd", "confirm", False), ("frame", "none"), ("cmd", "confirm", False), ("frame", (False, 0.9, (0, 0, 0, 0)))]  # success false
    ctx = PyContext(W, H)
    trace = []
    for step in script:
        if step[0] == "cmd":
            ctx.cmd(step[1], step[2])
            out = None
        else:
            u = step[1]
            out = ctx.frame(None if u in ("none", "err") else u)  # "none"/"err": update() (if called at all) returns Err
        trace.append({"step": [step[0], step[1] if step[0] == "cmd" else (step[1] if isinstance(step[1], str) else [bool(step[1][0]), step[1][1], list(step[1][2])]),
                               step[2] if step[0] == "cmd" else None],
                      "returned": out, "state": ctx.name(), "score": float(np.float32(ctx.score)), "bbox": ctx.bbox,
                      "selection": [ctx.sel.cursor_x, ctx.sel.cursor_y, ctx.sel.start_x, ctx.sel.start_y, ctx.sel.phase],
                      "sel_bbox": ctx.sel.bbox(), "lost": ctx.lost})
    # TimingStats trace (src/timing_stats.rs)
    iv = [int(v) for v in (synth.hash_u64(5, 300) % np.uint64(40000))]
    cv = [int(v) for v in (synth.hash_u64(6, 300) % np.uint64(9000))]
    tv = [int(v) for v in (synth.hash_u64(7, 300) % np.uint64(20000))]
    tstat = []
    for n in [0, 1, 5, 119, 120, 121, 300]:
        a, b, c = iv[:n][-120:], cv[:n][-120:], tv[:n][-120:]
        fps = 0.0 if not a or sum(a) == 0 else 1_000_000.0 / (sum(a) / len(a))
        tstat.append({"n": n, "fps": fps, "conv_ms": (sum(b) / len(b) / 1000.0) if b else 0.0, "track_ms": (sum(c) / len(c) / 1000.0) if c else 0.0})
    json.dump({"w": W, "h": H, "trace": trace, "timing": {"intervals": iv, "conv": cv, "track": tv, "checks": tstat}},
              open(os.path.join(GOLD, "state_traces.json"), "w"))


# ---------------------------------------------------------------------------------------------
def gen_resize():
    import cv2
    cases = []
    for k, c in enumerate([1, 2, 3, 5, 17, 40, 63, 64, 77, 100, 127, 128, 129, 200, 255, 256, 257, 300, 511, 512, 555, 700, 1000]):
        img = synth.hash_u8(900 + k, (c, c, 3))
        for dst in (128, 256):
            cases.append({"seed": 900 + k, "src": c, "dst": dst, "sha256": sha(cv2.resize(img, (dst, dst), interpolation=cv2.INTER_LINEAR))})
    json.dump({"source": f"cv2.resize INTER_LINEAR, OpenCV {cv2.__version__}", "cases": cases}, open(os.path.join(GOLD, "resize_golden.json"), "w"))


# ---------------------------------------------------------------------------------------------
def py_nv12_frame_to_rgb(nv12, w, h):
    p = nv12.astype(np.int32)
    y = p[: w * h].reshape(h, w)
    uv = p[w * h: w * h + (h // 2) * w].reshape(h // 2, w)
    u = np.repeat(np.repeat(uv[:, 0::2], 2, axis=0), 2, axis=1)
    v = np.repeat(np.repeat(uv[:, 1::2], 2, axis=0), 2, axis=1)
    yv = 298 * (y - 16)
    r = (yv + 409 * (v - 128) + 128) >> 8
    g = (yv - 100 * (u - 128) - 208 * (v - 128) + 128) >> 8
    b = (yv + 516 * (u - 128) + 128) >> 8
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def doctored_std():
    s = np.array([0.229, 0.224, 0.225])
    n = 1.0 / np.sum(1.0 / s ** 2)
    return (n / s[0], -n / s[1], -n / s[2], 0.0)


def gen_trackervit(model="nano"):
    """trackervit_<model>.json.  `nano` (D=64, L=2) is the quick fixture; `tiny` (D=192, L=12, the bench / headline model) pins the
    oracle and the GPU path at the depth and width where rounding accumulates (the headline model had no third-party check)."""
    import cv2
    from torch_model import export_onnx
    import tempfile

    tmp = tempfile.mkdtemp()
    out = {"source": f"cv2.TrackerVit (OpenCV {cv2.__version__}, DNN CPU), ONNX export of the same VTW1 file (opset 17)",
           "threshold": 0.2, "models": {}}
    for variant in ("stable", "wild"):
        wpath = weights.ensure_weight_file(model, tmp, variant=variant)
        whash = hashlib.sha256(open(wpath, "rb").read()).hexdigest()
        onnx = os.path.join(tmp, f"{model}_{variant}.onnx")
        export_onnx(wpath, onnx)

        def make():
            prm = cv2.TrackerVit_Params()
            prm.net, prm.stdvalue, prm.tracking_score_threshold = onnx, doctored_std(), 0.2
            return cv2.TrackerVit_create(prm)
        entry = {"weights_sha256": whash, "sequences": [], "single_steps": []}
        # (a) free-running sequences
        seqs = [("cfg1", synth.CONFIGS["cfg1"], 40), ("corner", synth.StreamSpec("corner", 640, 360, 31, [(2, 4, 90, 70, -3, -2)]), 30),
                ("small", synth.StreamSpec("small", 320, 240, 32, [(150, 100, 30, 24, 2, 1)]), 30)]
        if model == "tiny":  # the bench workload itself (cfg2, 1080p) leads
            seqs = [("cfg2", synth.CONFIGS["cfg2"], 40)] + [(n, s, 20) for n, s, _ in seqs]
        for name, spec, n in seqs:
            st = synth.SyntheticStream(spec)
            trk = make()
            f0 = py_nv12_frame_to_rgb(st.frame(0), spec.width, spec.height)
            box = st.target_boxes(0)[0]
            trk.init(f0, box)
            frames = []
            for i in range(n):
                ok, bb = trk.update(py_nv12_frame_to_rgb(st.frame(i), spec.width, spec.height))
                frames.append({"ok": bool(ok), "bbox": [int(v) for v in bb], "score": float(trk.getTrackingScore())})
            entry["sequences"].append({"name": name, "spec": {"w": spec.width, "h": spec.height, "seed": spec.seed, "targets": [list(t) for t in spec.targets]},
                                       "init_box": list(box), "frames": frames})
        # (b) independent single steps with boxes partly outside the frame
        spec = synth.StreamSpec("steps", 480, 270, 33, [(200, 100, 60, 50, 1, 1)])
        st = synth.SyntheticStream(spec)
        rgb0, rgb1 = [py_nv12_frame_to_rgb(st.frame(i), spec.width, spec.height) for i in (0, 5)]
        hv = synth.hash_u64(55, 200)
        k = 0
        for i in range(40):
            bw, bh = 20 + int(hv[k] % np.uint64(140)), 20 + int(hv[k + 1] % np.uint64(110))
            bx, by = int(hv[k + 2] % np.uint64(480 + 60)) - 60, int(hv[k + 3] % np.uint64(270 + 50)) - 50
            k += 4
            trk = make()
            try:
                trk.init(rgb0, (bx, by, bw, bh))
                ok, bb = trk.update(rgb1)
                entry["single_steps"].append({"box": [bx, by, bw, bh], "ok": bool(ok), "bbox": [int(v) for v in bb], "score": float(trk.getTrackingScore())})
            except cv2.error:
                entry["single_steps"].append({"box": [bx, by, bw, bh], "error": True})
        entry["steps_spec"] = {"w": spec.width, "h": spec.height, "seed": spec.seed, "targets": [list(t) for t in spec.targets], "frames": [0, 5]}
        out["models"][variant] = entry
    json.dump(out, open(os.path.join(GOLD, f"trackervit_{model}.json"), "w"))


def quirk_norm():
    """cv2 4.13 TrackerVit with its DEFAULT stdvalue normalises with OpenCV's quaternion Scalar division: the blob is
    (u8/255 - mean_c) * k_c with k = (s0, -s1, -s2) / (s0^2 + s1^2 + s2^2).  As the App. A.7 custom map blob = u8*scale_c + bias_c."""
    s = np.array([0.229, 0.224, 0.225])
    mean = np.array([0.485, 0.456, 0.406])
    k = np.array([s[0], -s[1], -s[2]]) / np.sum(s ** 2)
    return [float(v) for v in k / 255.0], [float(v) for v in -mean * k]


def gen_trackervit_variants():
    """trackervit_variants.json — the App. A.7 normalisation switch pinned against the third party: cv2.TrackerVit run with its default
    (undoctored) stdvalue, i.e. with the Scalar-division quirk, must equal the tracker configured with norm = quirk_norm().  (The other
    A.7 switches restate older OpenCV releases from recollection; no executable copy exists offline, they are checked GPU-vs-oracle only.)"""
    import cv2
    from torch_model import export_onnx
    import tempfile

    tmp = tempfile.mkdtemp()
    wpath = weights.ensure_weight_file("nano", tmp, variant="wild")
    onnx = os.path.join(tmp, "nano_wild.onnx")
    export_onnx(wpath, onnx)
    scale, bias = quirk_norm()
    out = {"source": f"cv2.TrackerVit (OpenCV {cv2.__version__}) with default stdvalue (Scalar-division quirk)", "model": "nano", "variant": "wild",
           "weights_sha256": hashlib.sha256(open(wpath, "rb").read()).hexdigest(), "norm_scale": scale, "norm_bias": bias, "sequences": []}
    for name, spec, n in [("cfg1", synth.CONFIGS["cfg1"], 20), ("corner", synth.StreamSpec("corner", 640, 360, 31, [(2, 4, 90, 70, -3, -2)]), 15)]:
        st = synth.SyntheticStream(spec)
        prm = cv2.TrackerVit_Params()
        prm.net, prm.tracking_score_threshold = onnx, 0.2
        trk = cv2.TrackerVit_create(prm)
        box = st.target_boxes(0)[0]
        trk.init(py_nv12_frame_to_rgb(st.frame(0), spec.width, spec.height), box)
        frames = []
        for i in range(n):
            ok, bb = trk.update(py_nv12_frame_to_rgb(st.frame(i), spec.width, spec.height))
            frames.append({"ok": bool(ok), "bbox": [int(v) for v in bb], "score": float(trk.getTrackingScore())})
        out["sequences"].append({"name": name, "spec": {"w": spec.width, "h": spec.height, "seed": spec.seed, "targets": [list(t) for t in spec.targets]},
                                 "init_box": list(box), "frames": frames})
    json.dump(out, open(os.path.join(GOLD, "trackervit_variants.json"), "w"))


def gen_yuy2():
    """yuy2_golden.json — (a) known answers and sha256 of frames converted by a pure-Python evaluation of the reference's BT.601
    integer formulas (/root/reference/src/nv12_convert.rs:24-30,124-126,41-43) on packed 4:2:2 input (rows of round_up_4(2*w) bytes,
    Y0 U Y1 V); (b) the maximum deviation from the third-party cv2.cvtColor(COLOR_YUV2RGB_YUY2), recorded for information
    (OpenCV uses 20-bit coefficients: a different rounding of the same BT.601 limited-range matrix);
    (c) sha256 of cv2.resize(INTER_LINEAR) up-scales of hash-generated RGB frames (640x512 -> 1280x1024 and odd cases)."""
    import cv2

    def clamp(v):
        return 0 if v < 0 else (255 if v > 255 else v)

    def px(y, u, v):
        yv = 298 * (y - 16)
        return (clamp((yv + 409 * (v - 128) + 128) >> 8), clamp((yv - 100 * (u - 128) - 208 * (v - 128) + 128) >> 8),
                clamp((yv + 516 * (u - 128) + 128) >> 8))

    def convert(buf, w, h):
        stride = (2 * w + 3) & ~3
        out = np.zeros((h, w, 3), np.uint8)
        if buf.size < stride * h:
            return out
        for r in range(h):
            row = buf[r * stride:(r + 1) * stride]
            for c in range(w):
                q = row[(c >> 1) * 4:(c >> 1) * 4 + 4]
                out[r, c] = px(int(q[(c & 1) * 2]), int(q[1]), int(q[3]))
        return out

    gold = {"kat": [], "frames": [], "resize": []}
    for (y, u, v) in [(16, 128, 128), (235, 128, 128), (81, 90, 240), (145, 54, 34), (41, 240, 110), (0, 0, 0), (255, 255, 255), (255, 0, 0), (0, 255, 255)]:
        gold["kat"].append({"yuv": [y, u, v], "rgb": list(px(y, u, v))})
    max_dev = 0
    for i, (w, h) in enumerate([(64, 48), (40, 6), (33, 5), (2, 2), (1, 3), (72, 10)]):
        stride = (2 * w + 3) & ~3
        buf = (synth.hash_u64(900 + i, stride * h) & np.uint64(0xFF)).astype(np.uint8)
        rgb = convert(buf, w, h)
        gold["frames"].append({"w": w, "h": h, "seed": 900 + i, "sha256": sha(rgb)})
        if w % 2 == 0:
            ref = cv2.cvtColor(buf.reshape(h, stride // 2, 2)[:, :w], cv2.COLOR_YUV2RGB_YUY2)
            max_dev = max(max_dev, int(np.abs(ref.astype(int) - rgb.astype(int)).max()))
    gold["max_abs_dev_from_cv2_cvtColor"] = max_dev
    for i, (sw, sh_, dw, dh) in enumerate([(640, 512, 1280, 1024), (64, 48, 128, 96), (33, 17, 100, 41), (640, 512, 960, 768), (50, 40, 25, 20)]):
        src = (synth.hash_u64(950 + i, sw * sh_ * 3) & np.uint64(0xFF)).astype(np.uint8).reshape(sh_, sw, 3)
        gold["resize"].append({"sw": sw, "sh": sh_, "dw": dw, "dh": dh, "seed": 950 + i,
                               "sha256": sha(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))})
    json.dump(gold, open(os.path.join(GOLD, "yuy2_golden.json"), "w"), indent=0)


def gen_keymap():
    """keymap.json — byte -> (UserCommand, fast) parsed out of the match in /root/reference/src/raw_mode_guard.rs:65-101."""
    src = open(os.path.join(REF, "raw_mode_guard.rs")).read()
    body = src[src.index("let cmd = match byte {"):src.index("if let Some(c) = cmd")]
    out = {}
    for m in re.finditer(r"^\s*((?:\d+\s*\|\s*)*\d+)\s*=>\s*(\{[^}]*\},?|[^,{]*,)", body, re.M | re.S):
        keys = [int(k) for k in re.findall(r"\d+", m.group(1))]
        rhs = m.group(2)
        c = re.search(r"UserCommand::(\w+)(?:\((true|false)\))?", rhs)
        for k in keys:
            out[str(k)] = [c.group(1), c.group(2) == "true"] if c else None
    assert out["91"] is None and out["113"] == ["Quit", False] and out["84"] == ["MoveUp", True] and len(out) == 33, (len(out), out)
    json.dump(out, open(os.path.join(GOLD, "keymap.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["kat", "glyphs", "overlay", "state", "resize", "trackervit", "trackervit_tiny", "variants", "keymap", "yuy2"]
    if "keymap" in which: gen_keymap()
    if "yuy2" in which: gen_yuy2()
    font = gen_glyphs() if ("glyphs" in which or "overlay" in which) else None
    if "kat" in which: gen_nv12_kat()
    if "overlay" in which: gen_overlay(font)
    if "state" in which: gen_state_traces()
    if "resize" in which: gen_resize()
    if "trackervit" in which: gen_trackervit("nano")
    if "trackervit_tiny" in which: gen_trackervit("tiny")
    if "variants" in which: gen_trackervit_variants()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
