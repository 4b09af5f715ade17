/* track_nv12.c — the reference's per-frame loop (src/pipeline.rs:67-184) against the C ABI, in plain C:
 *   convert + VitTrack::update + box overlay happen inside vt_tracker_update, in place on the caller's NV12 frame.
 *
 *   gcc -std=c99 -Iinclude examples/track_nv12.c -Lgstreamer_vit_tracker_b200 -lvittrack_b200 -o track_nv12
 *   ./track_nv12 model.vtw frames.nv12 1920 1080  880 480 160 120
 *
 * frames.nv12 = tightly packed NV12 frames (w*h*3/2 bytes each); the last four numbers are the target box in frame 0
 * (what the reference's SelectionState::to_bbox hands to VitTrack::init, src/tracker_context.rs:85-88).
 * Needs a B200: there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>

#include "vt_tracker.h"

static int fail(const char* what) {
    fprintf(stderr, "%s: %s\n", what, vt_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc != 9) {
        fprintf(stderr, "usage: %s model.vtw frames.nv12 width height x y w h\n", argv[0]);
        return 2;
    }
    int32_t shape[5];
    if (vt_weights_probe(argv[1], shape) != VT_OK) return fail("model file");
    printf("model: D=%d depth=%d heads=%d hidden=%d head_ch=%d\n", shape[0], shape[1], shape[2], shape[3], shape[4]);

    vt_config cfg;
    vt_config_default(&cfg);
    cfg.weights_path = argv[1];
    cfg.format = VT_FMT_NV12;
    cfg.width = atoi(argv[3]), cfg.height = atoi(argv[4]);
    cfg.box_overlay = 1;   /* draw_rect_nv12 + draw_crosshair_nv12 on the device, mirrored into the pinned frame */
    cfg.upload_window = 1; /* only the search window travels over PCIe */
    const size_t frame_bytes = (size_t)cfg.width * (size_t)cfg.height * 3 / 2;

    vt_tracker* trk = NULL;
    if (vt_tracker_create(&cfg, &trk) != VT_OK) return fail("vt_tracker_create");
    void* pinned = NULL; /* what a GstAllocator for the upstream element would hand out */
    if (vt_alloc_pinned(frame_bytes, &pinned) != VT_OK) return fail("vt_alloc_pinned");
    uint8_t* frame = (uint8_t*)pinned;

    FILE* f = fopen(argv[2], "rb");
    if (!f) {
        perror(argv[2]);
        return 1;
    }
    vt_bbox box;
    box.x = atoi(argv[5]), box.y = atoi(argv[6]), box.width = atoi(argv[7]), box.height = atoi(argv[8]);
    long n = 0;
    while (fread(frame, 1, frame_bytes, f) == frame_bytes) {
        if (n == 0) {
            if (vt_tracker_init(trk, 0, frame, frame_bytes, box) != VT_OK) return fail("vt_tracker_init");
        } else {
            vt_result r;
            if (vt_tracker_update(trk, frame, frame_bytes, &r) != VT_OK || r.status != VT_OK) return fail("vt_tracker_update");
            printf("frame %ld: %s score %.4f box (%d, %d, %d, %d)\n", n, r.success ? "ok  " : "lost", r.score, r.bbox.x, r.bbox.y, r.bbox.width,
                   r.bbox.height);
        }
        ++n;
    }
    fclose(f);
    vt_timing tm;
    if (vt_timing_get(trk, &tm) == VT_OK)
        printf("%ld frames, %.0f frames/s (host intervals), ViT stage %.3f ms, %.1f kernel launches per frame\n", n, tm.fps, tm.avg_vit_ms,
               tm.frames ? (double)tm.kernel_launches / (double)tm.frames : 0.0);
    vt_free_pinned(pinned);
    vt_tracker_destroy(trk);
    return 0;
}
