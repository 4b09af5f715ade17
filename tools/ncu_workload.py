#!/usr/bin/env python
"""Small, fixed workloads for ncu captures (a few frames each; numbers printed under ncu are never bench values).

  python tools/ncu_workload.py cfg2     1080p NV12, one target, tiny: 4 warm-up frames + 2 frames (device-resident, synchronous)
  python tools/ncu_workload.py cfg4     2160p NV12, 16 targets in one batched forward: 3 + 2 frames
  python tools/ncu_workload.py probe    vt_probe_frame on a pinned 1080p frame: 4 + 2 frames (HUD list, one synchronisation per frame)
  python tools/ncu_workload.py pixels   NV12->RGB (32 x 1080p), YUY2->RGB (256 x 640x512), RGB up-scale (64 x 640x512 -> 1280x1024): 2 launches each
  python tools/ncu_workload.py streams  8 concurrent 1080p streams (plain kernel forms), 3 + 2 frames each, driven round-robin from one thread

Typical use on the GPU box (B200_PROFILING.md):
  python tools/ncu_workload.py cfg4 > gpurun_out/plain.log 2>&1 && \\
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_workload.py cfg4
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gstreamer_vit_tracker_b200 import api, synth, weights  # noqa: E402

WDIR = os.path.join(tempfile.gettempdir(), "vt_b200_weights")


def tracker_frames(cfg, warm, frames, model="tiny"):
    spec = synth.CONFIGS[cfg]
    st = synth.SyntheticStream(spec)
    nt = len(spec.targets)
    trk = api.VitTrack.new(weights.ensure_weight_file(model, WDIR), spec.width, spec.height, fmt=spec.fmt, gemm_mode=1, box_overlay=True, max_targets=nt)
    host = [np.asarray(st.frame(i)).reshape(-1) for i in range(warm + frames)]
    dev = [torch.from_numpy(f).cuda() for f in host]
    for k, b in enumerate(st.target_boxes(0)):
        trk.init(host[0], api.BBox(*b), target=k)
    torch.cuda.synchronize()
    for i in range(warm + frames):
        r = trk.update_device(dev[i].data_ptr(), host[i].size)
    print(cfg, "last:", r[0], "launches", trk.timing().kernel_launches)


def probe_frames(warm, frames):
    spec = synth.CONFIGS["cfg2"]
    st = synth.SyntheticStream(spec)
    ctx = api.TrackerContext.new(weights.ensure_weight_file("tiny", WDIR), spec.width, spec.height, fmt="nv12", upload_window=True)
    pin = api.PinnedBuffer(st.frame_bytes())
    U = api.UserCommand
    x, y, w, h = st.target_boxes(0)[0]
    for _ in range((spec.width // 2 - x) // 10):
        ctx.handle_command(U.MoveLeft)
    for _ in range((spec.height // 2 - y) // 10):
        ctx.handle_command(U.MoveUp)
    ctx.handle_command(U.Confirm)
    pin.array[:] = st.frame(0)
    ctx.probe(pin.array)
    for _ in range(w // 10):
        ctx.handle_command(U.MoveRight)
    for _ in range(h // 10):
        ctx.handle_command(U.MoveDown)
    ctx.handle_command(U.Confirm)
    for i in range(warm + frames):
        pin.array[:] = st.frame(i)
        ctx.probe(pin.array)
    print("probe state", ctx.state_name(), "score", ctx.current_score)


def pixel_launches():
    trk = api.VitTrack.new(weights.ensure_weight_file("nano", WDIR), 1920, 1080, fmt="nv12")
    W, H, n = 1920, 1080, 32   # 100 MB in + 199 MB out per launch: more than the 126 MB L2, so the writes reach DRAM inside the launch
    src = torch.randint(0, 256, (n, W * H * 3 // 2), dtype=torch.uint8, device="cuda")
    dst = torch.empty((n, W * H * 3), dtype=torch.uint8, device="cuda")
    w, h, n2 = 640, 512, 256
    src2 = torch.randint(0, 256, (n2, w * h * 2), dtype=torch.uint8, device="cuda")
    dst2 = torch.empty((n2, w * h * 3), dtype=torch.uint8, device="cuda")
    n3 = 64
    up = torch.empty(n3 * 1280 * 1024 * 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        trk.nv12_to_rgb_device(src.data_ptr(), W * H * 3 // 2, dst.data_ptr(), W * H * 3, n)
        trk.yuy2_to_rgb_device(src2.data_ptr(), w * h * 2, dst2.data_ptr(), w * h * 3, w, h, n2)
        trk.resize_rgb_device_batch(dst2.data_ptr(), w * h * 3, w, h, up.data_ptr(), 1280 * 1024 * 3, 1280, 1024, n3)
    trk.sync()
    print("pixels ok: bytes per launch", n * W * H * 9 // 2, n2 * w * h * 5, n3 * (w * h * 3 + 1280 * 1024 * 3))


def stream_frames(n_streams, warm, frames):
    wpath = weights.ensure_weight_file("tiny", WDIR)
    trks, devs = [], []
    for i in range(n_streams):
        spec = synth.cfg5_stream(i)
        st = synth.SyntheticStream(spec)
        t = api.VitTrack.new(wpath, spec.width, spec.height, gemm_mode=1, box_overlay=True)
        fr = [st.frame(k) for k in range(warm + frames)]
        t.init(fr[0], api.BBox(*st.target_boxes(0)[0]))
        trks.append(t)
        devs.append([torch.from_numpy(f).cuda() for f in fr])
    torch.cuda.synchronize()
    for k in range(warm + frames):
        for i, t in enumerate(trks):
            t.submit_device(devs[i][k].data_ptr(), devs[i][k].numel())
        for t in trks:
            r = t.wait()
    print("streams last:", r[0])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    if what == "cfg2":
        tracker_frames("cfg2", 4, 2)
    elif what == "cfg4":
        tracker_frames("cfg4", 3, 2)
    elif what == "probe":
        probe_frames(4, 2)
    elif what == "pixels":
        pixel_launches()
    elif what == "streams":
        stream_frames(8, 3, 2)
    else:
        raise SystemExit(__doc__)
