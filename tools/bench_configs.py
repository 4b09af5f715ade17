#!/usr/bin/env python
"""Times the other BASELINE.json configurations (parity-test cases, not bench lines) on one GPU and prints one JSON line each:
cfg1 720p NV12, cfg3 640x512 RGB24 (the path the reference's main() runs, /root/reference/src/main.rs:49), cfg4 3840x2160 NV12 with
16 targets through one batched forward.  Device-timed stage breakdown from the tracker's own stamps; frames in pinned host memory
(end-to-end vt_tracker_update).  GPU only."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gstreamer_vit_tracker_b200 import api, synth, weights  # noqa: E402


def run(name, model="tiny", frames=200, warm=20, ring=16):
    name, _, limit = name.partition(":")  # "cfg4:8" = the first 8 targets of cfg4
    spec = synth.CONFIGS[name]
    if limit:
        import dataclasses
        spec = dataclasses.replace(spec, targets=spec.targets[:int(limit)])
    st = synth.SyntheticStream(spec)
    wpath = weights.ensure_weight_file(model, os.path.join(tempfile.gettempdir(), "vt_b200_weights"))
    nt = len(spec.targets)
    trk = api.VitTrack.new(wpath, spec.width, spec.height, fmt=spec.fmt, gemm_mode=1, box_overlay=True, max_targets=nt)
    fb = st.frame_bytes()
    pin = api.PinnedBuffer(ring * fb)
    host = pin.array.reshape(ring, fb)
    for i in range(ring):
        host[i] = np.asarray(st.frame(i)).reshape(-1)
    pristine = host.copy()
    for k, box in enumerate(st.target_boxes(0)):
        trk.init(host[0], api.BBox(*box), target=k)
    lat = []
    for i in range(warm + frames):
        j = i % ring
        if j == 0:
            host[:] = pristine
        t0 = time.perf_counter()
        res = trk.update_all(host[j])
        if i >= warm:
            lat.append(time.perf_counter() - t0)
    tm = trk.timing()
    lat = np.array(lat) * 1e3
    ok = sum(1 for r in res if r.success)
    print(json.dumps({
        "config": name, "resolution": f"{spec.width}x{spec.height}", "format": spec.fmt, "targets": nt, "model": model,
        "frames_per_s": 1e3 / lat.mean(), "target_frames_per_s": nt * 1e3 / lat.mean(), "p50_latency_ms": float(np.percentile(lat, 50)),
        "stages_ms": {k: getattr(tm, "avg_" + k) for k in ("h2d_ms", "preprocess_ms", "vit_ms", "decode_ms", "overlay_ms", "d2h_ms", "total_ms")},
        "targets_tracked_last_frame": ok}))


if __name__ == "__main__":
    for cfg in (sys.argv[1:] or ["cfg1", "cfg3", "cfg4"]):
        run(cfg)
