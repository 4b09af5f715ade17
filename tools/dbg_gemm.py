import sys, os, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo')
from gstreamer_vit_tracker_b200 import _lib
L = _lib.lib()
M, N, K = int(os.environ.get("MM", 320)), 192, 192
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K), dtype=np.float32); W = rng.standard_normal((N, K), dtype=np.float32) * 0.1
bias = rng.standard_normal(N, dtype=np.float32)
out = np.zeros((M, N), np.float32); err = C.c_int32(0)
f = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
st = L.vt_debug_gemm(0, M, N, K, f(A), f(W), f(bias), 3, 0, f(out), C.byref(err))
ref = A.astype(np.float64) @ W.astype(np.float64).T + bias
print(os.environ.get("VT_DBG_PERIOD"), os.environ.get("VT_DBG_RESID"), "status", st, "err", err.value, "maxdiff", float(np.abs(out - ref).max()), L.vt_last_error().decode() if st else "")
