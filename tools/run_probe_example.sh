#!/bin/bash
# Builds examples/probe_loop.c and runs it on 12 synthetic 720p frames: SELECT START -> SELECT END -> TRACKING (needs a B200).
set -e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
python - <<'PY'
import numpy as np, os, tempfile
from gstreamer_vit_tracker_b200 import synth, weights
spec = synth.CONFIGS["cfg1"]; st = synth.SyntheticStream(spec)
w = weights.ensure_weight_file("tiny", "/tmp/vt_b200_weights")
with open("/tmp/frames.nv12", "wb") as f:
    for i in range(12): f.write(np.ascontiguousarray(st.frame(i)).tobytes())
print(w)
PY
gcc -std=c99 -O1 -Iinclude examples/probe_loop.c -Lgstreamer_vit_tracker_b200 -lvittrack_b200 -Wl,-rpath,$PWD/gstreamer_vit_tracker_b200 -o /tmp/probe_loop
/tmp/probe_loop /tmp/vt_b200_weights/$(ls /tmp/vt_b200_weights | grep tiny | head -1) /tmp/frames.nv12 1280 720 "AAWW DDSS " > /tmp/out.nv12
ls -la /tmp/out.nv12
