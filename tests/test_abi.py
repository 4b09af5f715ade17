"""The C-ABI library loads on a CPU-only box and exports every symbol include/vt_tracker.h declares.
No compute calls: without a GPU the device entry points must fail loudly, never fall back."""
import ctypes as C
import os
import re

from conftest import ROOT


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "vt_tracker.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(built):
    from gstreamer_vit_tracker_b200 import _lib

    L = C.CDLL(_lib.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/vt_tracker.h but not exported"
    # and the Python binding table covers the same set
    assert set(_lib.SYMBOLS) == set(names)
    assert _lib.lib().vt_abi_version() == 2


def test_struct_layouts_match_header(built):
    from gstreamer_vit_tracker_b200 import _lib

    assert C.sizeof(_lib.vt_bbox) == 16
    assert C.sizeof(_lib.vt_result) == 32
    assert C.sizeof(_lib.vt_overlay_cmd) == 24 + 4 + 48
    cfg = _lib.vt_config()
    _lib.lib().vt_config_default(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(_lib.vt_config)
    assert (cfg.width, cfg.height, cfg.max_targets) == (1920, 1080, 1)
    assert abs(cfg.overlay_gate - 0.25) < 1e-7 and abs(cfg.score_threshold - 0.20) < 1e-7


def test_no_cpu_fallback(built):
    """Without a CUDA device create() must fail with VT_ERR_CUDA (or VT_ERR_WEIGHTS for a bad path on a GPU box)."""
    import torch

    from gstreamer_vit_tracker_b200 import _lib

    cfg = _lib.vt_config()
    _lib.lib().vt_config_default(C.byref(cfg))
    cfg.weights_path = b"/nonexistent/weights.vtw"
    h = C.c_void_p()
    st = _lib.lib().vt_tracker_create(C.byref(cfg), C.byref(h))
    assert st == (_lib.VT_ERR_WEIGHTS if torch.cuda.is_available() else _lib.VT_ERR_CUDA)
    assert not h.value
    assert _lib.lib().vt_last_error()


def test_product_does_not_touch_the_oracle():
    """Nothing under the product package or include/ may import, link or call oracle/."""
    pkg = os.path.join(ROOT, "gstreamer_vit_tracker_b200")
    for base, _, files in os.walk(pkg):
        if "build" in base.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "vt_oracle" not in txt and "vto_" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(base, f)


def test_null_arguments_are_rejected_not_dereferenced(built):
    """Every entry point called with NULL handles / buffers and zero sizes: a negative vt_status (or a harmless default), never a
    crash — the boundary is `extern "C"`, a null from the host language must not take the process down."""
    from gstreamer_vit_tracker_b200 import _lib

    lib = _lib.lib()
    for name, (res, args) in _lib.SYMBOLS.items():
        vals = []
        for a in args:
            if a in (C.c_int32, C.c_int, C.c_size_t, C.c_uint8, C.c_float, C.c_uint64, C.c_double):
                vals.append(a(0))
            elif a is _lib.vt_bbox:
                vals.append(_lib.vt_bbox(0, 0, 0, 0))
            else:
                vals.append(None)
        r = getattr(lib, name)(*vals)
        if name == "vt_timing_stats_create":  # no arguments: a real object
            assert r
            lib.vt_timing_stats_destroy(C.c_void_p(r))
        elif res is C.c_int32 and name not in ("vt_abi_version", "vt_command_from_key", "vt_tracker_model_dim", "vt_context_current_bbox",
                                                "vt_context_lost_frames"):
            assert r < 0, (name, r)


def test_header_is_plain_c_and_links(built, tmp_path):
    """include/vt_tracker.h is the FFI contract: it must compile as C99 (and C++11) on its own, and a C program must link against the
    library and get the documented defaults back — no torch, no C++ types in the boundary."""
    import shutil
    import subprocess

    from gstreamer_vit_tracker_b200 import _lib

    if not shutil.which("gcc"):
        import pytest
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "vt_tracker.h"\n#include <stdio.h>\n'
                   "int main(void) { vt_config c; vt_config_default(&c); int32_t cmd = -1, fast = -1;\n"
                   "  printf(\"%d %d %d %d %d\\n\", vt_abi_version(), (int)c.struct_size == (int)sizeof(vt_config), c.width, c.gemm_mode,\n"
                   "         vt_command_from_key((uint8_t)'q', &cmd, &fast));\n  return 0; }\n")
    inc, libdir = os.path.join(ROOT, "include"), os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, str(src), "-L", libdir, "-lvittrack_b200",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[:4] == ["2", "1", "1920", "1"], out
    if shutil.which("g++"):
        subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True,
                       capture_output=True)
    # the C examples (per-frame loop, interactive probe loop) build warning-free as well
    for ex in ("track_nv12", "probe_loop", "stream_group"):
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, os.path.join(ROOT, "examples", ex + ".c"), "-L", libdir,
                        "-lvittrack_b200", f"-Wl,-rpath,{libdir}", "-o", str(tmp_path / ex)], check=True, capture_output=True)
