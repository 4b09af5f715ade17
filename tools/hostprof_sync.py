import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.getcwd())
os.environ["VT_B200_HOSTPROF"] = "1"
from gstreamer_vit_tracker_b200 import api, synth, weights
spec = synth.CONFIGS["cfg2"]; st = synth.SyntheticStream(spec)
w = weights.ensure_weight_file("tiny", os.path.join(tempfile.gettempdir(), "vt_b200_weights"))
trk = api.VitTrack.new(w, spec.width, spec.height, gemm_mode=1, box_overlay=True, upload_window=True)
fb = st.frame_bytes(); ring = 32
pin = api.PinnedBuffer(ring * fb); host = pin.array.reshape(ring, fb)
for i in range(ring): host[i] = np.asarray(st.frame(i)).reshape(-1)
trk.init(host[0], api.BBox(*st.target_boxes(0)[0]))
lat = []
for i in range(400):
    t0 = time.perf_counter(); trk.update_all(host[i % ring]); lat.append(time.perf_counter() - t0)
lat = np.array(lat[50:]) * 1e6
print("sync p50 us", np.percentile(lat, 50), "mean", lat.mean())
tm = trk.timing(); print({k: getattr(tm, "avg_" + k) for k in ("h2d_ms", "preprocess_ms", "vit_ms", "decode_ms", "overlay_ms", "d2h_ms", "total_ms")})
trk.close()
