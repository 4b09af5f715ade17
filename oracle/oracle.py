"""ctypes binding of the CPU oracle (oracle/libvt_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvt_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("vt_oracle.c", "vt_oracle.h", "Makefile")]
    stale = not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return _SO


class BBox(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]

    def tuple(self) -> Tuple[int, int, int, int]:
        return (self.x, self.y, self.width, self.height)


class Result(C.Structure):
    _fields_ = [("success", C.c_int32), ("score", C.c_float), ("bbox", BBox)]


class Variant(C.Structure):
    """SURVEY.md App. A.7 switches (vto_variant)."""
    _fields_ = [("pad_plus1", C.c_int32), ("decode_window", C.c_int32), ("window", C.c_int32), ("norm_custom", C.c_int32),
                ("scale", C.c_float * 3), ("bias", C.c_float * 3)]


_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.vto_nv12_to_rgb.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, _u8p, C.c_int]
        L.vto_nv12_to_rgb.restype = None
        L.vto_yuy2_to_rgb.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.c_size_t, _u8p, C.c_int]
        L.vto_yuy2_to_rgb.restype = None
        L.vto_yuy2_stride.argtypes = [C.c_size_t]
        L.vto_yuy2_stride.restype = C.c_size_t
        for name, n_int in (("vto_draw_rect_nv12", 8), ("vto_draw_crosshair_nv12", 6), ("vto_draw_background_nv12", 7),
                            ("vto_draw_cursor_nv12", 4), ("vto_draw_selection_nv12", 7)):
            getattr(L, name).argtypes = [_u8p] + [C.c_int] * n_int
            getattr(L, name).restype = None
        L.vto_draw_text_nv12.argtypes = [_u8p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vto_draw_text_nv12.restype = None
        L.vto_get_glyph.argtypes = [C.c_int, _u8p]
        L.vto_get_glyph.restype = C.c_int
        for name, n_int in (("vto_draw_background_rgb", 6), ("vto_draw_rect_rgb", 10), ("vto_draw_crosshair_rgb", 8),
                            ("vto_draw_cursor_rgb", 4), ("vto_draw_selection_rgb", 7)):
            getattr(L, name).argtypes = [_u8p, C.c_size_t] + [C.c_int] * n_int
            getattr(L, name).restype = None
        L.vto_draw_text_rgb.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vto_draw_text_rgb.restype = None
        L.vto_timing_new.restype = C.c_void_p
        L.vto_timing_free.argtypes = [C.c_void_p]
        L.vto_timing_add_interval.argtypes = [C.c_void_p, C.c_uint64]
        L.vto_timing_add_times.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        for n in ("vto_timing_fps", "vto_timing_avg_conv_ms", "vto_timing_avg_track_ms"):
            getattr(L, n).argtypes = [C.c_void_p]
            getattr(L, n).restype = C.c_double
        L.vto_tracker_new.argtypes = [C.c_char_p, C.c_int]
        L.vto_tracker_new.restype = C.c_void_p
        L.vto_tracker_free.argtypes = [C.c_void_p]
        L.vto_tracker_set_threshold.argtypes = [C.c_void_p, C.c_float]
        L.vto_tracker_set_variant.argtypes = [C.c_void_p, C.POINTER(Variant)]
        L.vto_tracker_init.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, BBox]
        L.vto_tracker_init.restype = C.c_int
        L.vto_tracker_update.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.POINTER(Result)]
        L.vto_tracker_update.restype = C.c_int
        L.vto_tracker_get_rect.argtypes = [C.c_void_p, C.POINTER(BBox)]
        L.vto_tracker_set_rect.argtypes = [C.c_void_p, BBox]
        L.vto_tracker_last_maps.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p]
        L.vto_tracker_last_blobs.argtypes = [C.c_void_p, _f32p, _f32p]
        L.vto_crop_square.argtypes = [_u8p, C.c_int, C.c_int, BBox, C.c_int, _u8p, C.POINTER(C.c_int)]
        L.vto_crop_square.restype = C.c_int
        L.vto_resize_linear_u8c3.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
        L.vto_normalize_chw.argtypes = [_u8p, C.c_int, _f32p]
        L.vto_net_forward.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.vto_net_debug_tokens.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.vto_model_dim.argtypes = [C.c_void_p, C.c_int]
        L.vto_model_dim.restype = C.c_int
        L.vto_context_new.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.vto_context_new.restype = C.c_void_p
        L.vto_context_free.argtypes = [C.c_void_p]
        L.vto_context_handle_command.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.vto_context_process_frame.argtypes = [C.c_void_p, _u8p, C.POINTER(Result), C.c_int, C.POINTER(BBox)]
        L.vto_context_process_frame.restype = C.c_int
        L.vto_context_state.argtypes = [C.c_void_p]
        L.vto_context_state.restype = C.c_int
        L.vto_context_state_name.argtypes = [C.c_void_p]
        L.vto_context_state_name.restype = C.c_char_p
        L.vto_context_score.argtypes = [C.c_void_p]
        L.vto_context_score.restype = C.c_float
        L.vto_context_bbox.argtypes = [C.c_void_p, C.POINTER(BBox)]
        L.vto_context_bbox.restype = C.c_int
        L.vto_context_selection.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.vto_context_lost_frames.argtypes = [C.c_void_p]
        L.vto_context_lost_frames.restype = C.c_uint64
        _lib = L
    return _lib


def _u8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u8p)


def _f32(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


# ---- pixel code ---------------------------------------------------------------------------
def nv12_to_rgb(nv12: np.ndarray, width: int, height: int, threads: int = 1) -> np.ndarray:
    out = np.empty((height, width, 3), np.uint8)
    lib().vto_nv12_to_rgb(_u8(nv12), nv12.size, width, height, _u8(out), threads)
    return out


def yuy2_stride(width: int) -> int:
    return lib().vto_yuy2_stride(width)


def yuy2_to_rgb(yuy2: np.ndarray, width: int, height: int, threads: int = 1) -> np.ndarray:
    """YUY2 (packed 4:2:2, rows of yuy2_stride(width) bytes) -> HWC RGB; SURVEY.md §8(f) row 1."""
    yuy2 = np.ascontiguousarray(yuy2, dtype=np.uint8).reshape(-1)
    out = np.empty((height, width, 3), np.uint8)
    lib().vto_yuy2_to_rgb(_u8(yuy2), yuy2.size, width, height, _u8(out), threads)
    return out


def draw_rect_nv12(d, w, h, x, y, bw, bh, thickness=3, brightness=255):
    lib().vto_draw_rect_nv12(_u8(d), w, h, x, y, bw, bh, thickness, brightness)


def draw_crosshair_nv12(d, w, h, cx, cy, size=15, brightness=255):
    lib().vto_draw_crosshair_nv12(_u8(d), w, h, cx, cy, size, brightness)


def draw_text_nv12(d, w, h, text: str, x, y, scale, brightness):
    lib().vto_draw_text_nv12(_u8(d), w, h, text.encode("latin-1"), x, y, scale, brightness)


def draw_background_nv12(d, w, h, x, y, bw, bh, darkness):
    lib().vto_draw_background_nv12(_u8(d), w, h, x, y, bw, bh, darkness)


def draw_cursor_nv12(d, w, h, x, y):
    lib().vto_draw_cursor_nv12(_u8(d), w, h, x, y)


def draw_selection_nv12(d, w, h, sx, sy, cx, cy, selecting_area=True):
    lib().vto_draw_selection_nv12(_u8(d), w, h, sx, sy, cx, cy, int(selecting_area))


def get_glyph(ch: str):
    g = np.zeros(7, np.uint8)
    rc = lib().vto_get_glyph(ord(ch), _u8(g))
    return None if rc != 0 else g


def draw_background_rgb(d, w, h, x, y, bw, bh):
    lib().vto_draw_background_rgb(_u8(d), d.size, w, h, x, y, bw, bh)


def draw_rect_rgb(d, w, h, x, y, rw, rh, thickness=3, rgb=(0, 255, 0)):
    lib().vto_draw_rect_rgb(_u8(d), d.size, w, h, x, y, rw, rh, thickness, *rgb)


def draw_crosshair_rgb(d, w, h, cx, cy, size=15, rgb=(0, 255, 0)):
    lib().vto_draw_crosshair_rgb(_u8(d), d.size, w, h, cx, cy, size, *rgb)


def draw_cursor_rgb(d, w, h, cx, cy):
    lib().vto_draw_cursor_rgb(_u8(d), d.size, w, h, cx, cy)


def draw_text_rgb(d, w, h, text: str, x, y, scale, luma):
    lib().vto_draw_text_rgb(_u8(d), d.size, w, h, text.encode("latin-1"), x, y, scale, luma)


def draw_selection_rgb(d, w, h, sx, sy, cx, cy, selecting_area=True):
    lib().vto_draw_selection_rgb(_u8(d), d.size, w, h, sx, sy, cx, cy, int(selecting_area))


# ---- TimingStats ----------------------------------------------------------------------------
class TimingStats:
    def __init__(self):
        self._h = lib().vto_timing_new()

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vto_timing_free(self._h)
            self._h = None

    def add_interval(self, us: int):
        lib().vto_timing_add_interval(self._h, us)

    def add_times(self, conv: int, track: int):
        lib().vto_timing_add_times(self._h, conv, track)

    def fps(self) -> float:
        return lib().vto_timing_fps(self._h)

    def avg_conv_ms(self) -> float:
        return lib().vto_timing_avg_conv_ms(self._h)

    def avg_track_ms(self) -> float:
        return lib().vto_timing_avg_track_ms(self._h)


# ---- VitTrack -------------------------------------------------------------------------------
class VitTrack:
    """OpenCV-TrackerVit-semantics tracker over RGB24 frames (HWC u8)."""

    def __init__(self, weight_path: str, threads: int = 1):
        self._h = lib().vto_tracker_new(weight_path.encode(), threads)
        if not self._h:
            raise RuntimeError(f"oracle: cannot load weights {weight_path}")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vto_tracker_free(self._h)
            self._h = None

    def dim(self, which: int) -> int:
        return lib().vto_model_dim(self._h, which)

    def set_threshold(self, s: float):
        lib().vto_tracker_set_threshold(self._h, s)

    def set_variant(self, pad_plus1: bool = False, decode_window: int = 0, window: int = 0, norm=None):
        """App. A.7 switches; norm = (scale[3], bias[3]) -> blob = u8*scale[c] + bias[c]."""
        v = Variant(int(pad_plus1), int(decode_window), int(window), int(norm is not None))
        if norm is not None:
            for k in range(3):
                v.scale[k], v.bias[k] = float(norm[0][k]), float(norm[1][k])
        lib().vto_tracker_set_variant(self._h, C.byref(v))

    def init(self, rgb: np.ndarray, box) -> int:
        h, w, _ = rgb.shape
        return lib().vto_tracker_init(self._h, _u8(rgb), w, h, BBox(*box))

    def update(self, rgb: np.ndarray):
        """-> (rc, success, score, (x, y, w, h))"""
        h, w, _ = rgb.shape
        r = Result()
        rc = lib().vto_tracker_update(self._h, _u8(rgb), w, h, C.byref(r))
        return rc, bool(r.success), float(r.score), r.bbox.tuple()

    @property
    def rect(self):
        b = BBox()
        lib().vto_tracker_get_rect(self._h, C.byref(b))
        return b.tuple()

    @rect.setter
    def rect(self, box):
        lib().vto_tracker_set_rect(self._h, BBox(*box))

    def last_maps(self):
        cw, sm, om, cr = (np.empty(256, np.float32), np.empty(512, np.float32), np.empty(512, np.float32), np.empty(256, np.float32))
        lib().vto_tracker_last_maps(self._h, _f32(cw), _f32(sm), _f32(om), _f32(cr))
        return cw, sm, om, cr

    def last_blobs(self):
        sb, tb = np.empty((3, 256, 256), np.float32), np.empty((3, 128, 128), np.float32)
        lib().vto_tracker_last_blobs(self._h, _f32(sb), _f32(tb))
        return sb, tb

    def net_forward(self, template_blob: np.ndarray, search_blob: np.ndarray):
        conf, size, off = np.empty(256, np.float32), np.empty(512, np.float32), np.empty(512, np.float32)
        lib().vto_net_forward(self._h, _f32(np.ascontiguousarray(template_blob, np.float32)),
                              _f32(np.ascontiguousarray(search_blob, np.float32)), _f32(conf), _f32(size), _f32(off))
        return conf, size, off

    def debug_tokens(self, which: int) -> np.ndarray:
        out = np.empty((320, self.dim(0)), np.float32)
        lib().vto_net_debug_tokens(self._h, which, _f32(out))
        return out


def crop_square(rgb: np.ndarray, box, factor: int):
    h, w, _ = rgb.shape
    c = C.c_int(0)
    rc = lib().vto_crop_square(_u8(rgb), w, h, BBox(*box), factor, None, C.byref(c))
    if rc != 0:
        return rc, None
    out = np.empty((c.value, c.value, 3), np.uint8)
    lib().vto_crop_square(_u8(rgb), w, h, BBox(*box), factor, _u8(out), None)
    return 0, out


def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw, _ = src.shape
    out = np.empty((dh, dw, 3), np.uint8)
    lib().vto_resize_linear_u8c3(_u8(np.ascontiguousarray(src)), sw, sh, _u8(out), dw, dh)
    return out


def normalize_chw(hwc: np.ndarray) -> np.ndarray:
    size = hwc.shape[0]
    out = np.empty((3, size, size), np.float32)
    lib().vto_normalize_chw(_u8(np.ascontiguousarray(hwc)), size, _f32(out))
    return out


# ---- TrackerContext -------------------------------------------------------------------------
CMD = {"up": 0, "down": 1, "left": 2, "right": 3, "confirm": 4, "cancel": 5, "quit": 6}


class TrackerContext:
    def __init__(self, tracker: Optional[VitTrack], width: int, height: int):
        self._tracker = tracker
        self._h = lib().vto_context_new(tracker._h if tracker else None, width, height)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vto_context_free(self._h)
            self._h = None

    def handle_command(self, cmd: str, fast: bool = False):
        lib().vto_context_handle_command(self._h, CMD[cmd], int(fast))

    def process_frame(self, rgb: Optional[np.ndarray], scripted=None, scripted_err: bool = False):
        r = Result()
        if scripted is not None:
            r.success, r.score, r.bbox = int(scripted[0]), float(scripted[1]), BBox(*scripted[2])
        b = BBox()
        got = lib().vto_context_process_frame(self._h, _u8(rgb) if rgb is not None else None, C.byref(r), int(scripted_err), C.byref(b))
        return b.tuple() if got else None

    def state_name(self) -> str:
        return lib().vto_context_state_name(self._h).decode()

    @property
    def current_score(self) -> float:
        return lib().vto_context_score(self._h)

    @property
    def current_bbox(self):
        b = BBox()
        return b.tuple() if lib().vto_context_bbox(self._h, C.byref(b)) else None

    @property
    def selection(self):
        a = (C.c_int32 * 5)()
        lib().vto_context_selection(self._h, a)
        return tuple(a)

    @property
    def lost_frames(self) -> int:
        return lib().vto_context_lost_frames(self._h)
