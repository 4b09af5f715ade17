#!/usr/bin/env python
"""Prints the device-side timeline of one frame's tensor-core kernel chain (diagnostics, GPU only).

    VT_B200_TRACE=1 python tools/chain_timeline.py [--model tiny] [--gemm 1] [--frames 20]

Each record is stamped with %globaltimer by thread 0 of CTA 0 at kernel entry, after griddepcontrol.wait and at exit.
Columns: kernel, entry, prologue (entry -> dependency wait satisfied), body (wait -> end), gap (previous end -> this wait done).
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VT_B200_TRACE", "1")
from gstreamer_vit_tracker_b200 import api, synth, weights  # noqa: E402

NAMES = {1: "patch", 2: "qkv", 3: "proj", 4: "fc1", 5: "fc2", 6: "head", 10: "attn"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--gemm", type=int, default=1)
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--targets", type=int, default=1)
    ap.add_argument("--events", action="store_true", help="print the free-form events (TraceRec::event) of the first traced kernels")
    a = ap.parse_args()
    spec = synth.CONFIGS["cfg2"]
    st = synth.SyntheticStream(spec)
    wpath = weights.ensure_weight_file(a.model, os.path.join(tempfile.gettempdir(), "vt_b200_weights"))
    trk = api.VitTrack.new(wpath, spec.width, spec.height, fmt="nv12", gemm_mode=a.gemm, box_overlay=True, max_targets=a.targets)
    f0 = st.frame(0)
    for k in range(a.targets):
        trk.init(f0, api.BBox(*st.target_boxes(0)[0]), target=k)
    for i in range(a.frames):
        trk.update_all(st.frame(i % 8).copy())
    trk.debug_trace()
    trk.update_all(st.frame(3).copy())
    rec = trk.debug_trace().astype(np.int64)
    tm = trk.timing()
    rec = rec[np.argsort(rec[:, 1], kind="stable")]
    ev = rec[rec[:, 0] >= 256]
    rec = rec[rec[:, 0] < 256]
    t0 = rec[0, 1]
    if a.events:  # events of the traced CTA, relative to the dependency wait of the kernel they belong to (first 12 kernels)
        for kid, te, tw, tend, *_ in rec[:12]:
            mine = ev[(ev[:, 1] >= te) & (ev[:, 1] <= tend) & (ev[:, 2] == kid)]
            if len(mine):
                print(f"events of {NAMES.get(int(kid), str(kid))} @ {(te - t0) / 1e3:.2f}: " + " ".join(f"{int(c) - 256}:{(t - tw) / 1e3:.2f}" for c, t, *_ in mine))
    prev_end = None
    agg = {}
    print(f"{'kernel':8s} {'entry':>8s} {'prolog':>7s} {'body':>7s} {'gap':>7s} | marks 4..7 relative to the dependency wait"
          f"   (us; vit stage {tm.vit_ms * 1e3:.1f} us, total {tm.total_ms * 1e3:.1f} us)")
    print("  gemm: m6 accumulator ready, m4 values final (FC1: partial tile staged), m7 main copies issued, m5 LN exchange complete (FC1: chained accumulator ready); attn: m4 S ready, m5 row max done, m6 chunk 0 handed to the MMA, m7 chunk 4 handed over")
    for kid, te, tw, tend, m4, m5, m6, m7 in rec:
        name = NAMES.get(int(kid), str(kid))
        gap = (tw - prev_end) / 1e3 if prev_end is not None else 0.0
        marks = " ".join(f"{(m - tw) / 1e3:6.2f}" if m else "     -" for m in (m4, m5, m6, m7))
        print(f"{name:8s} {(te - t0) / 1e3:8.2f} {(tw - te) / 1e3:7.2f} {(tend - tw) / 1e3:7.2f} {gap:7.2f} | {marks}")
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += (tend - tw) / 1e3
        d[2] += gap
        prev_end = tend
    print("\nper kind: n, mean body us, mean gap-before us")
    for k, (n, b, g) in agg.items():
        print(f"  {k:6s} {n:3d} {b / n:7.2f} {g / n:7.2f}")
    print(f"chain span {(rec[-1, 3] - rec[0, 1]) / 1e3:.1f} us over {len(rec)} traced kernels")


if __name__ == "__main__":
    main()
